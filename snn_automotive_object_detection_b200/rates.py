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
K_s = sum_{n=0}^{T-1-s} kappa_n,  kappa_n = 0.9^{n+1} - 0.8^{n+1}, and its mean over the outputs only needs the
column sums of W.  Every statistic is therefore ONE pass of the library's leaky-integrator readout kernels over the
spike-train words with another per-step weight table (C ABI snn_li_readout_nhwc / snn_li_readout_rows): 1 for spike
counts, K_s / T for the time-mean membranes.  No spike train is ever expanded to per-step planes on the host side.
"""
import ctypes
from typing import List, Optional

import torch

from . import _lib
from .heads import RPNHeadSNN, FastRCNNPredictorSNNFull


def _tables(T: int):
    """(ones[32], K[32] / T) as ctypes double arrays; K_s = sum_{n <= T-1-s} kappa_n (float64)."""
    n = torch.arange(1, T + 1, dtype=torch.float64)
    cum = torch.cumsum(0.9 ** n - 0.8 ** n, 0)
    ones = (ctypes.c_double * 32)(*([1.0] * T + [0.0] * (32 - T)))
    mean_mem = (ctypes.c_double * 32)(*([float(cum[T - 1 - s]) / T for s in range(T)] + [0.0] * (32 - T)))
    return ones, mean_mem


def _ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _readout_nhwc(trains: torch.Tensor, table, w_a: torch.Tensor, w_b: torch.Tensor):
    """trains [N,H,W,C] words; w_a [n_a,C], w_b [n_b,C] -> ([N,n_a,H*W], [N,n_b,H*W])."""
    N, H, W, C = trains.shape
    dev = trains.device
    out_a = torch.empty(N, w_a.shape[0], H * W, device=dev, dtype=torch.float32)
    out_b = torch.empty(N, w_b.shape[0], H * W, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().snn_li_readout_nhwc(_ptr(trains), trains.element_size(), N, H * W, C, table, _ptr(w_a), w_a.shape[0],
                                             _ptr(w_b), w_b.shape[0], _ptr(out_a), _ptr(out_b), _stream(dev))
    _lib.check(rc, "snn_li_readout_nhwc")
    return out_a, out_b


def _readout_rows(trains: torch.Tensor, table, w_a: torch.Tensor, w_b: torch.Tensor):
    """trains [R,Hd] words; w_a [n_a,Hd], w_b [n_b,Hd] -> ([R,n_a], [R,n_b])."""
    R, Hd = trains.shape
    dev = trains.device
    out_a = torch.empty(R, w_a.shape[0], device=dev, dtype=torch.float32)
    out_b = torch.empty(R, w_b.shape[0], device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().snn_li_readout_rows(_ptr(trains), trains.element_size(), R, Hd, table, _ptr(w_a), w_a.shape[0],
                                             _ptr(w_b), w_b.shape[0], _ptr(out_a), _ptr(out_b), _stream(dev))
    _lib.check(rc, "snn_li_readout_rows")
    return out_a, out_b


def _require_trains(head):
    if head.last_spike_trains is None:
        raise RuntimeError("set head.record_spikes = True and run forward first")
    t0 = head.last_spike_trains[0]
    if not t0.is_cuda:
        raise RuntimeError("spike-rate report: expected the CUDA spike trains of a forward (B200); there is no CPU fallback")


@torch.no_grad()
def rpn_spike_rates_and_flops(head: RPNHeadSNN) -> List[torch.Tensor]:
    _require_trains(head)
    T, C, A = int(head.num_steps), head.in_channels, head.num_anchors
    ones, mean_mem = _tables(T)
    out = []
    dev = head.last_spike_trains[0].device
    # the mean over a readout's outputs only needs the column sums of its weight
    w_cls = head.conv_cls.weight.detach().float().view(A, C).sum(dim=0, keepdim=True).contiguous()
    w_box = head.conv_bbox.weight.detach().float().view(4 * A, C).sum(dim=0, keepdim=True).contiguous()
    one_row = torch.ones(1, C, device=dev, dtype=torch.float32)
    for l, trains in enumerate(head.last_spike_trains):                # [N,H,W,C] words
        N, H, W, _ = trains.shape
        if head.last_spike_counts is not None:                          # the conv epilogue's own per-image counts
            cnt = head.last_spike_counts[l].double().view(N, 1)
        else:
            cnt = _readout_nhwc(trains, ones, one_row, one_row)[0].double().sum(dim=(1, 2)).view(N, 1)
        r_s = (cnt / T / float(C * H * W)).float()
        m_o, m_b = _readout_nhwc(trains, mean_mem, w_cls, w_box)       # per pixel: sum over outputs of the time-mean membrane
        r_o = (m_o.double().sum(dim=(1, 2)) / float(A * H * W)).float().view(N, 1)
        r_b = (m_b.double().sum(dim=(1, 2)) / float(4 * A * H * W)).float().view(N, 1)
        f_s = torch.full((N, 1), float(9 * H * W * C * C), device=dev)
        f_o = torch.full((N, 1), float(H * W * C * A * 4), device=dev)       # sic: the reference's constants
        f_b = torch.full((N, 1), float(H * W * C * A), device=dev)
        out += [torch.hstack((r_s, f_s)), torch.hstack((r_o, f_o)), torch.hstack((r_b, f_b))]
    return out


@torch.no_grad()
def box_spike_rates_and_flops(head: FastRCNNPredictorSNNFull) -> List[torch.Tensor]:
    _require_trains(head)
    T, K, Hd, C = int(head.num_steps), head.in_channels, head.representation_size, head.num_classes
    t6, t7 = head.last_spike_trains
    R = t6.shape[0]
    dev = t6.device
    ones, mean_mem = _tables(T)
    nb = head.bbox_pred.out_features
    one_row = torch.ones(1, Hd, device=dev, dtype=torch.float32)
    if head.last_spike_counts is not None:                              # the readout kernel's own per-RoI counts
        c6, c7 = head.last_spike_counts[0].float().view(R, 1), head.last_spike_counts[1].float().view(R, 1)
    else:
        c6 = _readout_rows(t6, ones, one_row, one_row)[0]
        c7 = _readout_rows(t7, ones, one_row, one_row)[0]
    r6, r7 = c6 / float(T * Hd), c7 / float(T * Hd)
    w_c = (head.cls_score.weight.detach().float().sum(dim=0, keepdim=True) / float(C)).contiguous()
    w_b = (head.bbox_pred.weight.detach().float().sum(dim=0, keepdim=True) / float(nb)).contiguous()
    rc, rb = _readout_rows(t7, mean_mem, w_c, w_b)
    full = lambda v: torch.full((R, 1), float(v), device=dev)
    fb = Hd * C if head.only_one_bbox else Hd * C * 4
    return [torch.hstack((r6, full(K * Hd))), torch.hstack((r7, full(Hd * Hd))),
            torch.hstack((rc, full(Hd * C))), torch.hstack((rb, full(fb)))]


def energy_report(rpn_rates: List[torch.Tensor], box_rates: Optional[List[torch.Tensor]], T_rpn: int, T_det: int,
                  e_mac: float = 4.6e-12, e_ac: float = 0.9e-12):
    """The reference's energy estimate, train.py:470-517, from its 19-tensor list layout: the spiking layers are entries
    {0,3,6,9,12} (shared_lif per FPN level) and {15,16} (lif6, lif7); per layer  mean rate x T = spikes per neuron,
    ANN energy = FLOPs x 4.6 pJ, SNN energy = spikes x FLOPs x 0.9 pJ; the detector's per-RoI FLOPs are multiplied by
    the number of RoIs per image (the reference's x1000).  Returns (snn / ann, per-layer list)."""
    layers = []
    for i in range(0, len(rpn_rates), 3):
        v = rpn_rates[i]
        layers.append((f"LVL_{i // 3}", float(v[:, 0].double().mean()) * T_rpn, float(v[0, 1])))
    if box_rates is not None:
        for name, v in zip(("FC6", "FC7"), box_rates[:2]):
            n_img = max(rpn_rates[0].shape[0], 1) if rpn_rates else 1
            layers.append((name, float(v[:, 0].double().mean()) * T_det, float(v[0, 1]) * (v.shape[0] / n_img)))
    ann = sum(f * e_mac for _, _, f in layers)
    snn = sum(s * f * e_ac for _, s, f in layers)
    return snn / ann, [{"layer": n, "spikes_per_neuron": s, "flops": f, "snn_over_ann": s * e_ac / e_mac} for n, s, f in layers]


def energy_ratio(rates_and_flops: List[torch.Tensor], num_steps: int, e_mac: float = 4.6, e_ac: float = 0.9) -> float:
    """SNN / ANN energy of the spiking layers: rate * T * FLOPs * 0.9 pJ vs FLOPs * 4.6 pJ (train.py:506-507)."""
    snn = sum(float((t[:, 0] * num_steps * t[:, 1] * e_ac).sum()) for t in rates_and_flops)
    ann = sum(float((t[:, 1] * e_mac).sum()) for t in rates_and_flops)
    return snn / ann
