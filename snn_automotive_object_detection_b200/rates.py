"""Spike-rate / FLOP report in the reference's own format, from the spike trains the kernels emit.

The reference obtains this report only after hand-editing its sources (the alternative forward
bodies kept in string literals at rpn.py:126-200 and faster_rcnn.py:520-618, README.md:80-91).
Here it is a supported output: set `record_spikes = True` on the head, run forward, call
`rpn_spike_rates_and_flops(head)` / `box_spike_rates_and_flops(head)`.

Layout: per FPN level three tensors [N, 2] = (mean rate, layer FLOPs) for
shared_lif / lif_obj / lif_bbox; for the box head four tensors [R, 2] for lif6 / lif7 / lif_cls /
lif_bbox.  As in the reference the leaky-integrator "rates" average MEMBRANE values over time
(rpn.py:164-165, faster_rcnn.py:559-560) and the obj/bbox FLOP constants carry the reference's
swapped `* 4` (rpn.py:181-188).

The LI membrane is linear in the spikes, so  sum_t mem_t = W . sum_s K_s spk_s  with
K_s = sum_{n=0}^{T-1-s} kappa_n,  kappa_n = 0.9^{n+1} - 0.8^{n+1}.
These are statistics computed with torch ops on the device, off the hot path.
"""
from typing import List

import torch
import torch.nn.functional as F

from .heads import RPNHeadSNN, FastRCNNPredictorSNNFull


def _cum_kappa(T: int, device) -> torch.Tensor:
    n = torch.arange(1, T + 1, dtype=torch.float64)
    kap = 0.9 ** n - 0.8 ** n
    cum = torch.cumsum(kap, 0)                          # cum[m] = sum_{n<=m} kappa_n
    return torch.stack([cum[T - 1 - s] for s in range(T)]).to(torch.float32).to(device)   # K_s


def _count_and_weighted(trains: torch.Tensor, T: int):
    """trains: integer spike-train words [...].  Returns (spike count [...], sum_s K_s spk_s [...])."""
    w = trains.to(torch.int64)
    K = _cum_kappa(T, trains.device)
    cnt = torch.zeros(trains.shape, dtype=torch.float32, device=trains.device)
    ws = torch.zeros_like(cnt)
    for t in range(T):
        bit = ((w >> t) & 1).to(torch.float32)
        cnt += bit
        ws += K[t] * bit
    return cnt, ws


@torch.no_grad()
def rpn_spike_rates_and_flops(head: RPNHeadSNN) -> List[torch.Tensor]:
    if head.last_spike_trains is None:
        raise RuntimeError("set head.record_spikes = True and run forward first")
    T, C, A = int(head.num_steps), head.in_channels, head.num_anchors
    out = []
    w_cls = head.conv_cls.weight.detach().float()
    w_box = head.conv_bbox.weight.detach().float()
    for trains in head.last_spike_trains:                # [N,H,W,C]
        N, H, W, _ = trains.shape
        cnt, ws = _count_and_weighted(trains, T)
        r_s = (cnt.flatten(1).sum(1, keepdim=True) / T) / float(C * H * W)
        ws = ws.view(N, H * W, C)                        # NHWC; plain fp32 matmul (no TF32 conv path)
        r_o = (torch.matmul(ws, w_cls.view(A, C).t()).flatten(1) / T).mean(dim=1, keepdim=True)
        r_b = (torch.matmul(ws, w_box.view(4 * A, C).t()).flatten(1) / T).mean(dim=1, keepdim=True)
        dev = trains.device
        f_s = torch.full((N, 1), float(9 * H * W * C * C), device=dev)
        f_o = torch.full((N, 1), float(H * W * C * A * 4), device=dev)       # sic: the reference's constants
        f_b = torch.full((N, 1), float(H * W * C * A), device=dev)
        out += [torch.hstack((r_s, f_s)), torch.hstack((r_o, f_o)), torch.hstack((r_b, f_b))]
    return out


@torch.no_grad()
def box_spike_rates_and_flops(head: FastRCNNPredictorSNNFull) -> List[torch.Tensor]:
    if head.last_spike_trains is None:
        raise RuntimeError("set head.record_spikes = True and run forward first")
    T, K, Hd, C = int(head.num_steps), head.in_channels, head.representation_size, head.num_classes
    t6, t7 = head.last_spike_trains
    R = t6.shape[0]
    c6, _ = _count_and_weighted(t6, T)
    c7, ws7 = _count_and_weighted(t7, T)
    r6 = (c6 / T).mean(dim=1, keepdim=True)
    r7 = (c7 / T).mean(dim=1, keepdim=True)
    rc = (F.linear(ws7, head.cls_score.weight.detach().float()) / T).mean(dim=1, keepdim=True)
    rb = (F.linear(ws7, head.bbox_pred.weight.detach().float()) / T).mean(dim=1, keepdim=True)
    dev = t6.device
    full = lambda v: torch.full((R, 1), float(v), device=dev)
    fb = Hd * C if head.only_one_bbox else Hd * C * 4
    return [torch.hstack((r6, full(K * Hd))), torch.hstack((r7, full(Hd * Hd))),
            torch.hstack((rc, full(Hd * C))), torch.hstack((rb, full(fb)))]


def energy_ratio(rates_and_flops: List[torch.Tensor], num_steps: int, e_mac: float = 4.6, e_ac: float = 0.9) -> float:
    """SNN / ANN energy of the spiking layers: rate * T * FLOPs * 0.9 pJ vs FLOPs * 4.6 pJ (train.py:506-507)."""
    snn = sum(float((t[:, 0] * num_steps * t[:, 1] * e_ac).sum()) for t in rates_and_flops)
    ann = sum(float((t[:, 1] * e_mac).sum()) for t in rates_and_flops)
    return snn / ann
