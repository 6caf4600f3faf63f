import sys, torch
sys.path.insert(0, '.')
from oracle import snn_oracle as O
from snn_automotive_object_detection_b200 import RPNHeadSNN, unpack_trains
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W = O.reference_weights(seed=T)
w = [W["shared_conv"], W["conv_cls"], W["conv_bbox"]]
gen = torch.Generator().manual_seed(100 + T)
feats = [torch.randn(2, 256, h, wd, generator=gen) for (h, wd) in [(13, 22), (7, 9), (1, 1)]]
m = RPNHeadSNN(256, 3, T, mode="fp32_exact")
with torch.no_grad():
    m.shared_conv.weight.copy_(w[0]); m.conv_cls.weight.copy_(w[1]); m.conv_bbox.weight.copy_(w[2])
m = m.cuda(); m.record_spikes = True
lo, bb = m([f.cuda() for f in feats]); torch.cuda.synchronize()
rlo, rbb, tr = O.rpn_head_forward(feats, *w, T, record=True)
for l in range(3):
    got = unpack_trains(m.last_spike_trains[l].permute(0, 3, 1, 2).contiguous().cpu(), T)
    ref = tr[l]["spk"]
    diff = (got != ref)
    print("level", l, "agree", 1 - diff.float().mean().item(), "got ones", got.sum().item(), "ref ones", ref.sum().item())
    d = diff.any(dim=0).any(dim=1)   # [N,H,W]
    idx = d.nonzero()
    print(" bad pixels:", idx.tolist()[:40], "count", len(idx))
    if len(idx):
        n, h, ww = idx[0].tolist()
        print(" channels bad at first:", diff[:, n, :, h, ww].any(dim=0).nonzero().flatten().tolist()[:20], diff[:, n, :, h, ww].sum().item())
