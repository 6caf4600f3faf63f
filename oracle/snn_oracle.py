"""ORACLE / TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.  PARITY UNPINNED.

CPU (torch fp32) restatement of the reference's spiking detection heads:

  * rpn_head_forward   follows /root/reference/rpn.py:84-121  (RPNHeadSNN.forward)
  * box_head_forward   follows /root/reference/faster_rcnn.py:470-516
                       (FastRCNNPredictorSNNFull.forward)
  * rpn_head_rates / box_head_rates follow the spike-rate variants kept in
    string literals at rpn.py:126-200 and faster_rcnn.py:520-618
  * encoder_step / lif_step / li_step restate Norse 0.0.7
    (functional/lif.py, functional/leaky_integrator.py)

The neuron arithmetic lives in the third-party package norse==0.0.7
(README.md:13), which is absent from /root/reference and not installable
offline.  "Parity unpinned": the reference has no tests/golden vectors for this
path; this file is pinned by (a) closed-form known answers
(tests/test_oracle_known_answers.py), (b) goldens produced by running the
UNMODIFIED reference modules over oracle/norse_shim (oracle/gen_golden.py,
tests/golden/), which cross-checks loop order / state handling but shares the
restated Norse equations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product path
(snn_automotive_object_detection_b200/) never does.

Besides the step-by-step port there is a second, independent "flat"
restatement (rpn_head_flat / box_head_flat) that uses the three exact
structural identities the CUDA path relies on:
 (a) encoder spikes are a pure function of (x, t), so all live steps can be
     contracted in one batched conv / matmul;
 (b) the LIF cell integrates the previous step's current, so the last
     first-layer contraction is dead (fc6 dead for t > T-3, fc7 for t > T-2);
 (c) the LI readout is linear: mem_{T-1} = W . sum_t kappa_{T-1-t} spk_t,
     kappa_n = 0.9^{n+1} - 0.8^{n+1}.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

DT = 0.001
# dt * tau_mem_inv and dt * tau_syn_inv as Norse evaluates them: python float
# times 0-dim fp32 tensor -> fp32 0-dim tensor (0.1f and 0.2f).
_K_MEM = DT * torch.as_tensor(1.0 / 1e-2)
_K_SYN = DT * torch.as_tensor(1.0 / 5e-3)
V_TH_ENC = 0.25   # rpn.py:58, faster_rcnn.py:444
V_TH_LIF = 0.1    # rpn.py:67, faster_rcnn.py:449,452


# --------------------------------------------------------------------------
# Norse 0.0.7 primitives (inference arithmetic only)
# --------------------------------------------------------------------------
def encoder_step(x: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """lif_current_encoder with v_leak = v_reset = 0, v_th = 0.25."""
    dv = _K_MEM * ((0.0 - v) + x)
    v = v + dv
    z = ((v - torch.as_tensor(V_TH_ENC)) > 0).to(x.dtype)
    v = v - z * (v - 0.0)
    return z, v


def lif_step(cur: torch.Tensor, v: torch.Tensor, i: torch.Tensor):
    """lif_feed_forward_step with v_leak = v_reset = 0, v_th = 0.1."""
    dv = _K_MEM * ((0.0 - v) + i)
    v_dec = v + dv
    di = (-_K_SYN) * i
    i_dec = i + di
    z = ((v_dec - torch.as_tensor(V_TH_LIF)) > 0).to(cur.dtype)
    v_new = (1 - z) * v_dec + z * 0.0
    i_new = i_dec + cur
    return z, v_new, i_new, v_dec


def li_step(cur: torch.Tensor, v: torch.Tensor, i: torch.Tensor):
    """li_feed_forward_step with v_leak = 0."""
    i_new = i + cur
    dv = _K_MEM * ((0.0 - v) + i_new)
    v_new = v + dv
    di = (-_K_SYN) * i_new
    i_dec = i_new + di
    return v_new, i_dec


def encoder_spikes(x: torch.Tensor, num_steps: int) -> List[torch.Tensor]:
    v = torch.zeros_like(x)
    out = []
    for _ in range(num_steps):
        z, v = encoder_step(x, v)
        out.append(z)
    return out


def li_kernel(num_steps: int) -> torch.Tensor:
    """kappa_n = 0.9^{n+1} - 0.8^{n+1}, the LI impulse response, in float64."""
    n = torch.arange(1, num_steps + 1, dtype=torch.float64)
    return 0.9 ** n - 0.8 ** n


# --------------------------------------------------------------------------
# Step-by-step port of the two forwards
# --------------------------------------------------------------------------
@torch.no_grad()
def rpn_head_forward(features: Sequence[torch.Tensor], w_shared: torch.Tensor, w_cls: torch.Tensor,
                     w_bbox: torch.Tensor, num_steps: int, record: bool = False):
    """rpn.py:84-121.  Returns (logits, bbox_reg[, trace]).

    trace (record=True): per level a dict with 'spk' [T,N,C,H,W] uint8 and
    'v_dec' [T,N,C,H,W] fp32 (the pre-threshold membrane of shared_lif).
    """
    logits, bbox_reg, trace = [], [], []
    for feature in features:
        v = torch.zeros_like(feature)                               # rpn.py:93
        s_v = s_i = o_v = o_i = b_v = b_i = None                    # rpn.py:96
        spk_rec, vdec_rec = [], []
        for _ in range(num_steps):                                  # rpn.py:98
            z, v = encoder_step(feature, v)                         # rpn.py:101
            cur = F.conv2d(z, w_shared, padding=1)                  # rpn.py:105
            if s_v is None:
                s_v, s_i = torch.zeros_like(cur), torch.zeros_like(cur)
            spk, s_v, s_i, v_dec = lif_step(cur, s_v, s_i)          # rpn.py:106
            cur_o = F.conv2d(spk, w_cls)                            # rpn.py:110
            if o_v is None:
                o_v, o_i = torch.zeros_like(cur_o), torch.zeros_like(cur_o)
            o_v, o_i = li_step(cur_o, o_v, o_i)                     # rpn.py:111
            cur_b = F.conv2d(spk, w_bbox)                           # rpn.py:114
            if b_v is None:
                b_v, b_i = torch.zeros_like(cur_b), torch.zeros_like(cur_b)
            b_v, b_i = li_step(cur_b, b_v, b_i)                     # rpn.py:115
            if record:
                spk_rec.append(spk.to(torch.uint8))
                vdec_rec.append(v_dec)
        logits.append(o_v)                                          # rpn.py:118
        bbox_reg.append(b_v)                                        # rpn.py:119
        if record:
            trace.append({"spk": torch.stack(spk_rec), "v_dec": torch.stack(vdec_rec)})
    if record:
        return logits, bbox_reg, trace
    return logits, bbox_reg


@torch.no_grad()
def box_head_forward(x: torch.Tensor, w6: torch.Tensor, w7: torch.Tensor, w_cls: torch.Tensor,
                     w_bbox: torch.Tensor, num_steps: int, record: bool = False):
    """faster_rcnn.py:470-516.  Returns (cls_logits, bbox_deltas[, trace])."""
    x = x.flatten(start_dim=1)                                      # faster_rcnn.py:474
    v = torch.zeros_like(x)                                         # faster_rcnn.py:484
    R = x.shape[0]
    z6v = torch.zeros(R, w6.shape[0]); z6i = torch.zeros_like(z6v)
    z7v = torch.zeros(R, w7.shape[0]); z7i = torch.zeros_like(z7v)
    cv = torch.zeros(R, w_cls.shape[0]); ci = torch.zeros_like(cv)
    bv = torch.zeros(R, w_bbox.shape[0]); bi = torch.zeros_like(bv)
    rec6, rec7, vd6, vd7 = [], [], [], []
    for _ in range(num_steps):                                      # faster_rcnn.py:492
        z, v = encoder_step(x, v)                                   # :494
        cur = F.linear(z, w6)                                       # :498
        s6, z6v, z6i, v6d = lif_step(cur, z6v, z6i)                 # :499
        cur = F.linear(s6, w7)                                      # :500
        s7, z7v, z7i, v7d = lif_step(cur, z7v, z7i)                 # :501
        cv, ci = li_step(F.linear(s7, w_cls), cv, ci)               # :505-506
        bv, bi = li_step(F.linear(s7, w_bbox), bv, bi)              # :509-510
        if record:
            rec6.append(s6.to(torch.uint8)); rec7.append(s7.to(torch.uint8))
            vd6.append(v6d); vd7.append(v7d)
    if record:
        return cv, bv, {"spk6": torch.stack(rec6), "spk7": torch.stack(rec7),
                        "v_dec6": torch.stack(vd6), "v_dec7": torch.stack(vd7)}
    return cv, bv


# --------------------------------------------------------------------------
# Spike-rate variants (rpn.py:126-200, faster_rcnn.py:520-618)
# --------------------------------------------------------------------------
@torch.no_grad()
def rpn_head_rates(features, w_shared, w_cls, w_bbox, num_steps: int, num_anchors: int):
    """List of 3 tensors [N,2] per level: (mean rate, FLOPs) for shared / obj / bbox.
    The obj/bbox 'rates' average MEMBRANE values (rpn.py:164-165) and the FLOP
    constants carry the reference's swapped '*4' (rpn.py:181-188) -- reproduced as-is."""
    out = []
    C = w_shared.shape[1]
    for feature in features:
        N, _, H, W = feature.shape
        v = torch.zeros_like(feature)
        s_v = s_i = torch.zeros(N, w_shared.shape[0], H, W)
        o_v = o_i = torch.zeros(N, w_cls.shape[0], H, W)
        b_v = b_i = torch.zeros(N, w_bbox.shape[0], H, W)
        acc_s = torch.zeros(N, w_shared.shape[0] * H * W)
        acc_o = torch.zeros(N, w_cls.shape[0] * H * W)
        acc_b = torch.zeros(N, w_bbox.shape[0] * H * W)
        for _ in range(num_steps):
            z, v = encoder_step(feature, v)
            spk, s_v, s_i, _ = lif_step(F.conv2d(z, w_shared, padding=1), s_v, s_i)
            o_v, o_i = li_step(F.conv2d(spk, w_cls), o_v, o_i)
            b_v, b_i = li_step(F.conv2d(spk, w_bbox), b_v, b_i)
            acc_s += spk.flatten(1); acc_o += o_v.flatten(1); acc_b += b_v.flatten(1)
        r_s = (acc_s / num_steps).mean(dim=1, keepdim=True)
        r_o = (acc_o / num_steps).mean(dim=1, keepdim=True)
        r_b = (acc_b / num_steps).mean(dim=1, keepdim=True)
        f_s = torch.full((N, 1), float(9 * H * W * C * C))
        f_o = torch.full((N, 1), float(H * W * C * num_anchors * 4))
        f_b = torch.full((N, 1), float(H * W * C * num_anchors))
        out += [torch.hstack((r_s, f_s)), torch.hstack((r_o, f_o)), torch.hstack((r_b, f_b))]
    return out


@torch.no_grad()
def box_head_rates(x, w6, w7, w_cls, w_bbox, num_steps: int, only_one_bbox: bool = False):
    """4 tensors [R,2]: (mean rate, FLOPs) for fc6 / fc7 / cls / bbox (faster_rcnn.py:597-603)."""
    x = x.flatten(1)
    R, K = x.shape
    Hd, C = w6.shape[0], w_cls.shape[0]
    v = torch.zeros_like(x)
    s6v = s6i = torch.zeros(R, Hd); s7v = s7i = torch.zeros(R, Hd)
    cv = ci = torch.zeros(R, C); bv = bi = torch.zeros(R, w_bbox.shape[0])
    a6 = torch.zeros(R, Hd); a7 = torch.zeros(R, Hd); ac = torch.zeros(R, C); ab = torch.zeros(R, w_bbox.shape[0])
    for _ in range(num_steps):
        z, v = encoder_step(x, v)
        s6, s6v, s6i, _ = lif_step(F.linear(z, w6), s6v, s6i)
        s7, s7v, s7i, _ = lif_step(F.linear(s6, w7), s7v, s7i)
        cv, ci = li_step(F.linear(s7, w_cls), cv, ci)
        bv, bi = li_step(F.linear(s7, w_bbox), bv, bi)
        a6 += s6; a7 += s7; ac += cv; ab += bv
    mk = lambda a: (a / num_steps).mean(dim=1, keepdim=True)
    f6 = torch.full((R, 1), float(K * Hd)); f7 = torch.full((R, 1), float(Hd * Hd))
    fc = torch.full((R, 1), float(Hd * C))
    fb = torch.full((R, 1), float(Hd * C if only_one_bbox else Hd * C * 4))
    return [torch.hstack((mk(a6), f6)), torch.hstack((mk(a7), f7)),
            torch.hstack((mk(ac), fc)), torch.hstack((mk(ab), fb))]


# --------------------------------------------------------------------------
# Independent "flat" restatement (the algebra the CUDA path uses)
# --------------------------------------------------------------------------
def _lif_unroll(cur: torch.Tensor, num_steps: int, t0: int = 0) -> torch.Tensor:
    """cur: [T_live, ...] currents injected at steps t0 .. t0+T_live-1.  Returns
    spikes [num_steps, ...] (uint8)."""
    v = torch.zeros_like(cur[0]); i = torch.zeros_like(cur[0])
    out = []
    for t in range(num_steps):
        k = t - t0
        c = cur[k] if 0 <= k < cur.shape[0] else torch.zeros_like(cur[0])
        z, v, i, _ = lif_step(c, v, i)
        out.append(z.to(torch.uint8))
    return torch.stack(out)


@torch.no_grad()
def rpn_head_flat(features, w_shared, w_cls, w_bbox, num_steps: int):
    T = num_steps
    kap = li_kernel(T).to(torch.float32)
    logits, bbox, spikes = [], [], []
    for feature in features:
        N = feature.shape[0]
        if T > 1:
            z = torch.cat(encoder_spikes(feature, T)[: T - 1], dim=0)       # (a) + (b)
            cur = F.conv2d(z, w_shared, padding=1).view(T - 1, N, w_shared.shape[0], *feature.shape[2:])
        else:
            cur = torch.zeros(1, N, w_shared.shape[0], *feature.shape[2:])
        spk = _lif_unroll(cur, T)                                            # [T,N,C,H,W]
        wsum = torch.zeros(N, w_shared.shape[0], *feature.shape[2:])
        for t in range(T):                                                   # (c)
            wsum += kap[T - 1 - t] * spk[t].to(torch.float32)
        logits.append(F.conv2d(wsum, w_cls)); bbox.append(F.conv2d(wsum, w_bbox)); spikes.append(spk)
    return logits, bbox, spikes


@torch.no_grad()
def box_head_flat(x, w6, w7, w_cls, w_bbox, num_steps: int):
    T = num_steps
    x = x.flatten(1); R = x.shape[0]
    kap = li_kernel(T).to(torch.float32)
    if T > 2:
        z = torch.stack(encoder_spikes(x, T)[: T - 2])                       # fc6 live for t <= T-3
        cur6 = F.linear(z.view(-1, x.shape[1]), w6).view(-1, R, w6.shape[0])
        spk6 = _lif_unroll(cur6, T)
        s6_live = spk6[1: T - 1].to(torch.float32)                           # fc7 live for 1 <= t <= T-2
        cur7 = F.linear(s6_live.reshape(-1, w6.shape[0]), w7).view(-1, R, w7.shape[0])
        spk7 = _lif_unroll(cur7, T, t0=1)
    else:
        spk6 = torch.zeros(T, R, w6.shape[0], dtype=torch.uint8)
        spk7 = torch.zeros(T, R, w7.shape[0], dtype=torch.uint8)
    wsum = torch.zeros(R, w7.shape[0])
    for t in range(T):
        wsum += kap[T - 1 - t] * spk7[t].to(torch.float32)
    return F.linear(wsum, w_cls), F.linear(wsum, w_bbox), spk6, spk7


# --------------------------------------------------------------------------
# Helpers shared by tests and the bench's CPU leg
# --------------------------------------------------------------------------
def pack_trains(spk: torch.Tensor) -> torch.Tensor:
    """[T, ...] {0,1} -> [...] int64 spike-train word, bit t = spike at step t
    (the layout the CUDA path emits: one word per neuron, time packed)."""
    T = spk.shape[0]
    w = torch.zeros(spk.shape[1:], dtype=torch.int64)
    for t in range(T):
        w |= spk[t].to(torch.int64) << t
    return w


CITYSCAPES_LEVELS = [(192, 384), (96, 192), (48, 96), (24, 48), (12, 24)]
BDD_LEVELS = [(192, 344), (96, 172), (48, 86), (24, 43), (12, 22)]


def synthetic_inputs(levels, n_images: int, rois_per_image: int = 1000, channels: int = 256,
                     pool: int = 7, seed: int = 1234) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """features randn level 0->4, then RoI features, one generator."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(n_images, channels, h, w, generator=g) for (h, w) in levels]
    rois = torch.randn(n_images * rois_per_image, channels, pool, pool, generator=g)
    return feats, rois


def reference_weights(in_channels: int = 256, num_anchors: int = 3, box_in: int = 12544,
                      rep: int = 1024, num_classes: int = 9, only_one_bbox: bool = False,
                      seed: int = 0) -> Dict[str, torch.Tensor]:
    """Weights drawn exactly as the reference constructors draw them after
    torch.manual_seed(seed): RPNHeadSNN creates shared_conv, conv_cls, conv_bbox
    (default Conv2d init) then re-draws each with normal_(std=0.01) in
    self.modules() order (rpn.py:65-82); FastRCNNPredictorSNNFull creates fc6,
    fc7, cls_score, bbox_pred with the default nn.Linear init (faster_rcnn.py:448-467)."""
    torch.manual_seed(seed)
    sc = torch.nn.Conv2d(in_channels, in_channels, 3, 1, 1, bias=False)
    cc = torch.nn.Conv2d(in_channels, num_anchors, 1, 1, bias=False)
    cb = torch.nn.Conv2d(in_channels, num_anchors * 4, 1, 1, bias=False)
    for layer in (sc, cc, cb):
        torch.nn.init.normal_(layer.weight, std=0.01)
    fc6 = torch.nn.Linear(box_in, rep, bias=False)
    fc7 = torch.nn.Linear(rep, rep, bias=False)
    cs = torch.nn.Linear(rep, num_classes, bias=False)
    bp = torch.nn.Linear(rep, 4 if only_one_bbox else num_classes * 4, bias=False)
    return {"shared_conv": sc.weight.detach(), "conv_cls": cc.weight.detach(), "conv_bbox": cb.weight.detach(),
            "fc6": fc6.weight.detach(), "fc7": fc7.weight.detach(),
            "cls_score": cs.weight.detach(), "bbox_pred": bp.weight.detach()}
