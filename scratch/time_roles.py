"""Profiling experiment (GPU box): where the MMA-issuing thread of each CTA pair waits, per spike-GEMM launch.
Usage: python scratch/time_roles.py [phase] [mode]   (phase 0 conv, 1 fc6 dual tiles, 3 fc6 single tiles, 2 fc7)"""
import sys
import torch
sys.path.insert(0, ".")
from snn_automotive_object_detection_b200 import _lib, RPNHeadSNN, FastRCNNPredictorSNNFull

phase = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
lib = _lib.load()
torch.manual_seed(0)
dev = "cuda"
shapes = [(192, 384), (96, 192), (48, 96), (24, 48), (12, 24)]
feats = [torch.randn(2, 256, h, w, device=dev) for h, w in shapes]
rois = torch.randn(2000, 256, 7, 7, device=dev)
rpn = RPNHeadSNN(256, 3, 8, mode=mode).to(dev).eval(); rpn.record_rates = True
box = FastRCNNPredictorSNNFull(12544, 1024, 9, 12, mode=mode).to(dev).eval(); box.record_rates = True
for _ in range(3):
    rpn(feats); box(rois)
torch.cuda.synchronize()
buf = torch.zeros(74, 12, dtype=torch.int64, device=dev)
lib.snn_set_role_timers(buf.data_ptr(), phase)
rpn(feats); box(rois)
torch.cuda.synchronize()
lib.snn_set_role_timers(None, -1)
b = buf.cpu().double()
tot = b[:, 0]
names = ["total", "acc_empty", "b_ready(local)", "b_peer", "a_full(weights)", "tiles", "epilogue role", "epi wait acc_full", "entry->first tile", "entry->exit", "entry->exit (ns)"]
print(f"phase {phase} mode {mode}: pairs with work {(tot > 0).sum().item()}")
for k, n in enumerate(names):
    col = b[:, k][tot > 0]
    frac = (col / tot[tot > 0]).mean().item() if k not in (0, 5) else float("nan")
    print(f"  {n:18s} mean {col.mean().item():12.0f}  min {col.min().item():12.0f}  max {col.max().item():12.0f}  mean share {frac:.3f}")

eff = (b[:, 9][tot > 0] / b[:, 10][tot > 0].clamp(min=1)).mean().item() * 1e3
print(f"  effective SM clock inside the kernel: {eff:.0f} MHz (cycles / globaltimer ns)")
