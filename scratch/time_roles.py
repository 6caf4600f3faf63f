"""Cycle accounting of the spike GEMM roles (debug counters written by producer warp 8 and the MMA thread)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from snn_automotive_object_detection_b200 import RPNHeadSNN, FastRCNNPredictorSNNFull, _lib
mode = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
lib = _lib.load()
lib.snn_debug_set_times.argtypes = [ctypes.c_void_p]
torch.manual_seed(0)
box = FastRCNNPredictorSNNFull(12544, 1024, 9, 12, mode=mode).cuda(); box.record_rates = True
rpn = RPNHeadSNN(256, 3, 8, mode=mode).cuda(); rpn.record_rates = True
g = torch.Generator().manual_seed(1)
x = torch.randn(2000, 256, 7, 7, generator=g).cuda()
feats = [torch.randn(2, 256, h, w, generator=g).cuda() for (h, w) in [(192, 384), (96, 192), (48, 96), (24, 48), (12, 24)]]
for _ in range(2):
    box(x); rpn(feats)
torch.cuda.synchronize()
names = ["prod: wait input", "prod: wait b_empty", "prod: expand", "prod: publish", "mma: wait acc_empty", "mma: wait b_ready", "mma: wait b_peer", "mma: total"]
for what, fn in (("box head (fc6 then fc7 overwrite: fc7 shown for CTAs it uses)", lambda: box(x)), ("rpn conv", lambda: rpn(feats))):
    t = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
    lib.snn_debug_set_times(ctypes.c_void_p(t.data_ptr()))
    fn(); torch.cuda.synchronize()
    lib.snn_debug_set_times(None)
    tt = t.cpu().double()
    lead, peer = tt[0::2], tt[1::2]
    print(what)
    for k, n in enumerate(names):
        print("   %-22s leader CTAs mean %10.0f   peer CTAs mean %10.0f" % (n, lead[:, k].mean().item(), peer[:, k].mean().item()))
