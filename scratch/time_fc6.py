"""Cycle accounting of the fc6-shaped spike GEMM alone (debug counters)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from snn_automotive_object_detection_b200 import _lib
from tests._util import vp, stream, prepared_fc
mode = {"fp16x2": 3, "bf16": 1, "fp32_exact": 0}[sys.argv[1] if len(sys.argv) > 1 else "fp16x2"]
lib = _lib.load()
lib.snn_debug_set_times.argtypes = [ctypes.c_void_p]
R, K, M, T, T_live = 2000, 12544, 1024, 12, 11
g = torch.Generator().manual_seed(0)
planes = (torch.rand(R, K // 8, 16, generator=g) < 0.02).to(torch.uint8).cuda() * 3
w = (torch.randn(M, K, generator=g) * 0.01).cuda()
wp = prepared_fc(w, mode)
trains = torch.zeros(R, M, dtype=torch.int16, device="cuda")
p6 = torch.zeros(R, M // 8, 16, dtype=torch.uint8, device="cuda")
names = ["prod: wait input", "prod: wait b_empty", "prod: expand", "prod: publish", "mma: wait acc_empty", "mma: wait b_ready", "mma: wait b_peer", "mma: total"]
for rep in range(3):
    t = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
    lib.snn_debug_set_times(ctypes.c_void_p(t.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.snn_fc_lif_layer(vp(planes), 2, 0, R, K, M, T, 0, T_live, mode, vp(wp), vp(trains), vp(p6), None, 0, stream())
    e1.record(); torch.cuda.synchronize()
    lib.snn_debug_set_times(None)
print("fc6 ms", e0.elapsed_time(e1))
tt = t.cpu().double(); lead, peer = tt[0::2], tt[1::2]
for k, n in enumerate(names):
    print("   %-22s leader mean %10.0f   peer mean %10.0f" % (n, lead[:, k].mean().item(), peer[:, k].mean().item()))
print("k-blocks per CTA pair ~", 400 / 74 * 196)
