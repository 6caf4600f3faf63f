#!/usr/bin/env python
"""Where the 0.57 ms of rpn_select_proposals go (GPU box): key kernel, the five int64 top-k calls, decode kernel."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snn_automotive_object_detection_b200 import _lib

LEVELS = [(192, 384), (96, 192), (48, 96), (24, 48), (12, 24)]
N, A = 2, 3
dev = torch.device("cuda")
torch.manual_seed(0)
obj = [torch.randn(N, A, h, w, device=dev) for (h, w) in LEVELS]
lib = _lib.load()


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


sizes = [o[0].numel() for o in obj]
keys_flat = torch.empty(N * sum(sizes), dtype=torch.int64, device=dev)
keys, off = [], 0
for n in sizes:
    keys.append(keys_flat[off:off + N * n].view(N, n)); off += N * n
VP, IA = ctypes.c_void_p * 5, ctypes.c_int * 5


def make_keys():
    lib.snn_rpn_topk_keys(VP(*[t.data_ptr() for t in obj]), IA(*[t.shape[2] for t in obj]), IA(*[t.shape[3] for t in obj]), 5, N, A,
                          VP(*[t.data_ptr() for t in keys]), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


make_keys()
out = {"keys_kernel_ms": timed(make_keys)}
for l, k in enumerate(keys):
    kk = min(1000, sizes[l])
    out[f"topk_int64_level{l}_ms"] = timed(lambda: k.topk(kk, dim=1))
    v = obj[l].view(N, -1)
    out[f"topk_fp32_level{l}_ms"] = timed(lambda: v.topk(kk, dim=1))
    out[f"sort_int64_level{l}_ms"] = timed(lambda: k.sort(dim=1, descending=True))
print(json.dumps(out))
