"""Drop-in replacements for the reference's two spiking heads, computed by the
sm_100a kernels in csrc/ through the C ABI (include/snn_heads.h).

  RPNHeadSNN                 mirrors /root/reference/rpn.py:33-121
  FastRCNNPredictorSNNFull   mirrors /root/reference/faster_rcnn.py:414-516

Same constructor arguments, parameter names / shapes (state_dict compatible:
shared_conv / conv_cls / conv_bbox and fc6 / fc7 / cls_score / bbox_pred, all
bias-free), same initialisation order, same forward signatures and outputs
(last-step leaky-integrator membranes).  Inference only: no autograd graph is
built, and a forward in training mode with gradients enabled raises instead of
handing detached outputs to the losses (the reference trains through these heads
with Norse's surrogate gradient, train.py:149-200 -- that path is out of scope).
CUDA tensors only -- there is no CPU fallback.
"""
import ctypes
from typing import List, NamedTuple, Optional, Tuple

import torch
from torch import nn, Tensor

from . import _lib


class EncoderParameters(NamedTuple):
    """Stand-in for norse LIFParameters as used for `p_enc` (rpn.py:58, faster_rcnn.py:444);
    callers only read `.v_th` (custom_utils.py:322,329)."""
    tau_syn_inv: Tensor = torch.as_tensor(1.0 / 5e-3)
    tau_mem_inv: Tensor = torch.as_tensor(1.0 / 1e-2)
    v_leak: Tensor = torch.as_tensor(0.0)
    v_th: Tensor = torch.as_tensor(0.25)
    v_reset: Tensor = torch.as_tensor(0.0)
    method: str = "super"
    alpha: Tensor = torch.as_tensor(100.0)


_TRAIN_DTYPE = {1: torch.uint8, 2: torch.int16, 4: torch.int32}


def _ptr(t: Optional[Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _PreparedWeight:
    """16-bit pieces of an fp32 weight, rebuilt only when the parameter changes.

    "Changes" is detected through (data_ptr, _version, mode, device, shape).  In-place edits through `.data`
    (weight.data.copy_, EMA / clipping on .data) do not bump `_version`: call the module's
    `invalidate_weight_cache()` after such an edit.  load_state_dict() and .to()/.cuda()/.float() invalidate
    by themselves (module hooks below)."""

    def __init__(self):
        self.key = None
        self.buf = None

    def invalidate(self):
        self.key = None
        self.buf = None

    def get(self, weight: Tensor, mode: int, conv: bool):
        key = (weight.data_ptr(), weight._version, mode, str(weight.device), tuple(weight.shape))
        if key != self.key:
            lib = _lib.load()
            w = weight.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            rows = w.shape[0]
            cols = w[0].numel()
            nbytes = lib.snn_prepared_weight_bytes(rows, cols, mode)
            buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
            with torch.cuda.device(w.device):
                if conv:
                    rc = lib.snn_prepare_conv3x3_weights(_ptr(w), rows, w.shape[1], mode, _ptr(buf), _stream(w.device))
                else:
                    rc = lib.snn_prepare_fc_weights(_ptr(w), rows, cols, mode, _ptr(buf), _stream(w.device))
            _lib.check(rc, "snn_prepare_weights")
            self.key, self.buf = key, buf
        return self.buf


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


class _InferenceOnly(nn.Module):
    """Shared host logic of the two heads: weight-cache invalidation hooks and the training-mode guard."""

    def _prepared(self):
        return [v for v in self.__dict__.values() if isinstance(v, _PreparedWeight)]

    def invalidate_weight_cache(self):
        """Forget the prepared (16-bit pieces) copies of the weights; the next forward rebuilds them."""
        for w in self._prepared():
            w.invalidate()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_weight_cache()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_weight_cache()
        return out

    def _refuse_training(self, inputs):
        if self.training and torch.is_grad_enabled() and (
                any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in inputs)):
            raise RuntimeError(
                f"{type(self).__name__} is inference-only: its CUDA kernels build no autograd graph, so in training "
                "mode the RPN / RoI losses would be computed from detached outputs.  Call model.eval() and/or run "
                "under torch.no_grad(); train with the reference's Norse heads and load the state_dict here "
                "(parameter names and shapes are identical).")


def _require_cuda(t: Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (B200); the spiking heads have no CPU fallback")


class EncodedRoIs(NamedTuple):
    """RoI features that are already the box head's encoder output (detection_post.FusedRoIAlignEncoder):
    `words` [R, K] spike-train words for num_steps - 1 encoder steps."""
    words: Tensor
    num_steps: int


class RPNHeadSNN(_InferenceOnly):
    """Spiking RPN head (reference: rpn.py:33-121).

    Args (as the reference): in_channels, num_anchors, num_steps.
    Extra keyword `mode`: "fp32_exact" (default; weights fed to the tensor cores as 3 bf16
    pieces -> parity mode), "bf16x2" or "bf16" (throughput mode).
    After a forward, `last_spike_trains` (if `record_spikes`) holds per level the
    shared_lif spike trains [N,H,W,C] (bit t = spike at step t) and `last_spike_counts`
    (if `record_rates`) the total spikes per level and image [levels, N].
    """

    _version = 2

    def __init__(self, in_channels: int, num_anchors: int, num_steps, mode="fp32_exact") -> None:
        super().__init__()
        self.num_steps = num_steps
        self.dt = 0.001
        self.p_enc = EncoderParameters()
        self.in_channels = in_channels
        self.num_anchors = num_anchors
        self.shared_conv = nn.Conv2d(in_channels, in_channels, kernel_size=(3, 3), stride=(1, 1), padding=1, bias=False)
        self.conv_cls = nn.Conv2d(in_channels, num_anchors, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.conv_bbox = nn.Conv2d(in_channels, num_anchors * 4, kernel_size=(1, 1), stride=(1, 1), bias=False)
        for layer in self.modules():                                   # rpn.py:78-82
            if isinstance(layer, nn.Conv2d):
                torch.nn.init.normal_(layer.weight, std=0.01)
        self.mode = mode
        self.record_spikes = False
        self.record_rates = False
        self.last_spike_trains: Optional[List[Tensor]] = None
        self.last_spike_counts: Optional[Tensor] = None
        self.last_launch_count = 0
        self._w_shared = _PreparedWeight()
        self._ws = _Workspace()

    def forward(self, x: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        self._refuse_training(x)
        with torch.no_grad():
            return self._forward(x)

    def _forward(self, x: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        lib = _lib.load()
        mode = _lib.mode_id(self.mode)
        T = int(self.num_steps)
        L = len(x)
        if L == 0:
            return [], []
        dev = x[0].device
        feats = []
        for f in x:
            _require_cuda(f, "RPNHeadSNN.forward")
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise RuntimeError(f"RPNHeadSNN.forward: expected [N,{self.in_channels},H,W], got {tuple(f.shape)}")
            feats.append(f.detach().float().contiguous())
        N = feats[0].shape[0]
        C, A = self.in_channels, self.num_anchors
        Hs = (ctypes.c_int * L)(*[f.shape[2] for f in feats])
        Ws = (ctypes.c_int * L)(*[f.shape[3] for f in feats])
        # one allocation, outputs back to back (all logits, then all box deltas): the library clears them with a
        # single memset before the two CTAs of a pair add their channel halves
        sizes = [N * A * f.shape[2] * f.shape[3] for f in feats]
        flat = torch.empty(5 * sum(sizes), device=dev, dtype=torch.float32)
        logits, bbox, off = [], [], 0
        for f, n in zip(feats, sizes):
            logits.append(flat[off:off + n].view(N, A, f.shape[2], f.shape[3])); off += n
        for f, n in zip(feats, sizes):
            bbox.append(flat[off:off + 4 * n].view(N, 4 * A, f.shape[2], f.shape[3])); off += 4 * n
        if N == 0:
            return logits, bbox
        w_prep = self._w_shared.get(self.shared_conv.weight, mode, conv=True)
        w_cls = self.conv_cls.weight.detach().float().contiguous()
        w_bbox = self.conv_bbox.weight.detach().float().contiguous()
        nbytes = lib.snn_rpn_head_workspace_bytes(Hs, Ws, L, N, C, T, mode)
        if nbytes == 0:
            _lib.check(-1, "snn_rpn_head_workspace_bytes")
        ws = self._ws.get(nbytes, dev)
        VP = ctypes.c_void_p * L
        trains, trains_arr = None, None
        if self.record_spikes:
            tdt = _TRAIN_DTYPE[lib.snn_train_word_bytes(T)]
            trains = [torch.empty(N, f.shape[2], f.shape[3], C, device=dev, dtype=tdt) for f in feats]
            trains_arr = VP(*[t.data_ptr() for t in trains])
        counts = torch.zeros(L, N, device=dev, dtype=torch.int64) if self.record_rates else None
        with torch.cuda.device(dev):
            rc = lib.snn_rpn_head_forward(
                VP(*[f.data_ptr() for f in feats]), Hs, Ws, L, N, C, A, T, mode, _ptr(w_prep), _ptr(w_cls), _ptr(w_bbox),
                VP(*[t.data_ptr() for t in logits]), VP(*[t.data_ptr() for t in bbox]),
                trains_arr, _ptr(counts), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "snn_rpn_head_forward")
        self.last_launch_count = lib.snn_last_launch_count()
        self.last_spike_trains, self.last_spike_counts = trains, counts
        return logits, bbox


class FastRCNNPredictorSNNFull(_InferenceOnly):
    """Spiking box head + predictor (reference: faster_rcnn.py:414-516).

    Args (as the reference): in_channels, representation_size, num_classes, num_steps,
    only_one_bbox=False.  Extra keyword `mode` as for RPNHeadSNN.
    """

    def __init__(self, in_channels, representation_size, num_classes, num_steps, only_one_bbox=False,
                 mode="fp32_exact"):
        super().__init__()
        self.num_steps = num_steps
        self.dt = 0.001
        self.in_channels = in_channels
        self.representation_size = representation_size
        self.num_classes = num_classes
        self.p_enc = EncoderParameters()
        self.fc6 = nn.Linear(in_channels, representation_size, bias=False)
        self.fc7 = nn.Linear(representation_size, representation_size, bias=False)
        self.cls_score = nn.Linear(representation_size, num_classes, bias=False)
        self.only_one_bbox = only_one_bbox
        self.bbox_pred = nn.Linear(representation_size, 4 if only_one_bbox else num_classes * 4, bias=False)
        self.mode = mode
        self.record_spikes = False
        self.record_rates = False
        self.last_spike_trains: Optional[Tuple[Tensor, Tensor]] = None
        self.last_spike_counts: Optional[Tensor] = None
        self.last_launch_count = 0
        self._w6 = _PreparedWeight()
        self._w7 = _PreparedWeight()
        self._ws = _Workspace()

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        self._refuse_training([x.words if isinstance(x, EncodedRoIs) else x])
        with torch.no_grad():
            return self._forward(x)

    def _forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        lib = _lib.load()
        mode = _lib.mode_id(self.mode)
        T = int(self.num_steps)
        encoded = isinstance(x, EncodedRoIs)
        if encoded:                      # RoIAlign + encoder already fused upstream
            if x.num_steps != T:
                raise RuntimeError(f"FastRCNNPredictorSNNFull.forward: RoIs were encoded for {x.num_steps} steps, head runs {T}")
            x = x.words
            _require_cuda(x, "FastRCNNPredictorSNNFull.forward")
            R, K = x.shape
        else:
            _require_cuda(x, "FastRCNNPredictorSNNFull.forward")
            x = x.detach().flatten(start_dim=1).float().contiguous()          # faster_rcnn.py:474
            R, K = x.shape
        if K != self.in_channels:
            raise RuntimeError(f"FastRCNNPredictorSNNFull.forward: expected {self.in_channels} features, got {K}")
        dev = x.device
        Hd, C = self.representation_size, self.num_classes
        nb = self.bbox_pred.out_features
        cls = torch.empty(R, C, device=dev, dtype=torch.float32)
        box = torch.empty(R, nb, device=dev, dtype=torch.float32)
        if R == 0:
            self.last_spike_trains, self.last_spike_counts = None, None
            return cls, box
        w6 = self._w6.get(self.fc6.weight, mode, conv=False)
        w7 = self._w7.get(self.fc7.weight, mode, conv=False)
        w_cls = self.cls_score.weight.detach().float().contiguous()
        w_box = self.bbox_pred.weight.detach().float().contiguous()
        nbytes = lib.snn_box_head_workspace_bytes(R, K, Hd, T, mode)
        if nbytes == 0:
            _lib.check(-1, "snn_box_head_workspace_bytes")
        ws = self._ws.get(nbytes, dev)
        tr6 = tr7 = None
        if self.record_spikes:
            tdt = _TRAIN_DTYPE[lib.snn_train_word_bytes(T)]
            tr6 = torch.empty(R, Hd, device=dev, dtype=tdt)
            tr7 = torch.empty(R, Hd, device=dev, dtype=tdt)
        counts = torch.zeros(2, R, device=dev, dtype=torch.int32) if self.record_rates else None
        with torch.cuda.device(dev):
            fwd = lib.snn_box_head_forward_encoded if encoded else lib.snn_box_head_forward
            rc = fwd(_ptr(x), R, K, Hd, C, nb, T, mode, _ptr(w6), _ptr(w7), _ptr(w_cls), _ptr(w_box),
                                          _ptr(cls), _ptr(box), _ptr(tr6), _ptr(tr7), _ptr(counts), _ptr(ws), ws.numel(),
                                          _stream(dev))
        _lib.check(rc, "snn_box_head_forward")
        self.last_launch_count = lib.snn_last_launch_count()
        self.last_spike_trains = (tr6, tr7) if self.record_spikes else None
        self.last_spike_counts = counts
        return cls, box


def unpack_trains(trains: Tensor, num_steps: int) -> Tensor:
    """[..] spike-train words -> [T, ..] uint8 spikes (bit t = spike at step t)."""
    w = trains.to(torch.int64)
    if trains.dtype == torch.int16:
        w = w & 0xFFFF
    elif trains.dtype == torch.int32:
        w = w & 0xFFFFFFFF
    return torch.stack([((w >> t) & 1).to(torch.uint8) for t in range(num_steps)])
