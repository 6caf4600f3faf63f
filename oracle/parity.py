"""ORACLE / TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.  PARITY UNPINNED (see snn_oracle.py).

The parity bar of BASELINE.json's north_star, stated once and used by tests/, smoke() and
bench.py's verify leg:

  * per-neuron spike agreement >= 99.9 % over (neuron, step);
  * every neuron whose train differs must, at its FIRST differing step, have the oracle's
    pre-threshold membrane within 1e-5 of the threshold (or sit downstream of such a neuron);
  * logits / deltas within 1e-3 of the tensor's scale -- EVERYWHERE.  A position fed by a flipped
    neuron is not exempted: a flip of neuron c moves output o by exactly
    w[o, c] * sum_t kappa_{T-1-t} (got_t - ref_t)  (the LI readout is linear in the spikes,
    Norse li_feed_forward_step; rpn.py:110-115, faster_rcnn.py:505-510), so the allowed deviation of
    an output is  rel * scale + sum_c |w[o, c]| * |d_kappa[c]|  with d_kappa computed from the two
    spike trains.  With no flip the second term is zero and the plain 1e-3 bar applies.
"""
from typing import Dict, Optional

import torch

from . import snn_oracle as O

BAND = 1e-5
REL = 1e-3


def unpack_trains(trains: torch.Tensor, num_steps: int) -> torch.Tensor:
    """[...] spike-train words (bit t = spike at step t) -> [T, ...] uint8."""
    w = trains.to(torch.int64)
    if trains.dtype == torch.int16:
        w = w & 0xFFFF
    elif trains.dtype == torch.int32:
        w = w & 0xFFFFFFFF
    return torch.stack([((w >> t) & 1).to(torch.uint8) for t in range(num_steps)])


def kappa_delta(got_spk: torch.Tensor, ref_spk: torch.Tensor) -> torch.Tensor:
    """[T, ...] spikes of both sides -> [...] fp32: sum_t kappa_{T-1-t} (got_t - ref_t), the change a
    train difference makes to the neuron's weight in the last leaky-integrator membrane."""
    T = got_spk.shape[0]
    kap = O.li_kernel(T)                                   # kappa_n, n = 0..T-1 (float64)
    d = torch.zeros(got_spk.shape[1:], dtype=torch.float64)
    for t in range(T):
        d += kap[T - 1 - t] * (got_spk[t].to(torch.float64) - ref_spk[t].to(torch.float64))
    return d.to(torch.float32)


def flip_stats(got_spk: torch.Tensor, ref_spk: torch.Tensor, ref_vdec: torch.Tensor, v_th: float = O.V_TH_LIF,
               upstream_rows: Optional[torch.Tensor] = None, band: float = BAND) -> Dict:
    """Spike agreement and the flips that the near-threshold rule does not explain.
    upstream_rows: bool [rows] -- rows (dim 0 of the neuron shape) that already carry a flip in the layer feeding
    this one; every neuron of such a row is downstream of a flip."""
    diff = got_spk != ref_spk
    out = {"neuron_steps": diff.numel(), "agreement": 1.0 - diff.float().mean().item(), "flipped_neurons": 0,
           "unexplained": 0, "max_first_flip_distance": 0.0}
    if not diff.any():
        return out
    anyd = diff.any(dim=0)
    first = diff.to(torch.uint8).argmax(dim=0)
    vd = torch.gather(ref_vdec, 0, first.unsqueeze(0)).squeeze(0)
    dist = (vd - v_th).abs()
    explained = dist < band
    if upstream_rows is not None:
        explained = explained | upstream_rows.view(-1, *([1] * (anyd.dim() - 1))).expand_as(anyd)
    own = anyd if upstream_rows is None else anyd & ~upstream_rows.view(-1, *([1] * (anyd.dim() - 1))).expand_as(anyd)
    out["flipped_neurons"] = int(anyd.sum())
    out["unexplained"] = int((anyd & ~explained).sum())
    out["max_first_flip_distance"] = float(dist[own].max()) if own.any() else 0.0
    return out


def bounded_close(got: torch.Tensor, ref: torch.Tensor, slack: torch.Tensor, rel: float = REL) -> Dict:
    """|got - ref| <= rel * max|ref| + slack, element-wise (slack >= 0, same shape, zero where no neuron flipped)."""
    got = got.detach().cpu().float(); ref = ref.detach().cpu().float()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs()
    excess = err - slack
    return {"ok": bool((excess <= rel * scale).all()), "scale": scale,
            "max_err_unflipped": float((err * (slack == 0)).max()), "max_excess": float(excess.max()),
            "positions_with_slack": int((slack > 0).sum()), "positions": slack.numel()}


def rpn_level_parity(lo, bb, trains_nchw, ref_lo, ref_bb, ref_trace, w_cls, w_bbox, T) -> Dict:
    """One FPN level.  trains_nchw: [N, C, H, W] spike-train words of shared_lif from the device (CPU tensor)."""
    got = unpack_trains(trains_nchw, T)
    st = flip_stats(got, ref_trace["spk"], ref_trace["v_dec"])
    dk = kappa_delta(got, ref_trace["spk"]).abs()                                   # [N, C, H, W]
    A = w_cls.shape[0]
    s_lo = torch.einsum("oc,nchw->nohw", w_cls.reshape(A, -1).abs(), dk)
    s_bb = torch.einsum("oc,nchw->nohw", w_bbox.reshape(4 * A, -1).abs(), dk)
    st["logits"] = bounded_close(lo, ref_lo, s_lo)
    st["bbox"] = bounded_close(bb, ref_bb, s_bb)
    st["pixels_with_flip"] = int((dk.amax(dim=1) > 0).sum())
    return st


def box_parity(cls, box, t6, t7, ref_cls, ref_box, ref_trace, w_cls, w_bbox, T) -> Dict:
    """t6 / t7: [R, Hd] spike-train words of lif6 / lif7 from the device (CPU tensors)."""
    g6, g7 = unpack_trains(t6, T), unpack_trains(t7, T)
    s6 = flip_stats(g6, ref_trace["spk6"], ref_trace["v_dec6"])
    rows6 = (g6 != ref_trace["spk6"]).any(dim=0).any(dim=1)
    s7 = flip_stats(g7, ref_trace["spk7"], ref_trace["v_dec7"], upstream_rows=rows6)
    dk = kappa_delta(g7, ref_trace["spk7"]).abs()                                    # [R, Hd]
    rows7 = dk.amax(dim=1) > 0
    return {"lif6": s6, "lif7": s7, "rows": int(rows6.numel()), "rows_with_flip": int((rows6 | rows7).sum()),
            "cls": bounded_close(cls, ref_cls, dk @ w_cls.abs().t()),
            "bbox": bounded_close(box, ref_box, dk @ w_bbox.abs().t())}


def assert_layer(st: Dict, what: str, min_agree: float = 0.999, strict: bool = True):
    assert st["agreement"] >= min_agree, f"{what}: spike agreement {st['agreement']}"
    if strict:
        assert st["unexplained"] == 0, (f"{what}: {st['unexplained']} flipped neurons outside the {BAND} band "
                                        f"(farthest first flip {st['max_first_flip_distance']:.3e})")


def assert_close(c: Dict, what: str):
    assert c["ok"], (f"{what}: exceeds 1e-3 of scale {c['scale']:.4g} beyond the flip bound by {c['max_excess']:.3e} "
                     f"(max err on unflipped positions {c['max_err_unflipped']:.3e})")
