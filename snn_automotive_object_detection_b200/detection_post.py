"""The two steps around the spiking heads that SURVEY.md section 8f ranks "next":

  * rank 1 -- RPN proposal selection fed from the head's native layout.  The reference
    (RegionProposalNetwork.forward, rpn.py:636-670; filter_proposals, rpn.py:448-525) permutes every level's
    logits/deltas to (H, W, A) order, generates all anchors, decodes ALL boxes and only then takes the
    per-level top-k.  `rpn_select_proposals` takes the top-k on the NCHW logits as the head wrote them and lets
    one small CUDA kernel (csrc/aux_kernels.cuh::rpn_decode_selected_kernel, through the C ABI) regenerate the
    anchors of the selected entries, gather their deltas, decode and apply the sigmoid.  NMS stays torchvision.
  * rank 4 -- `postprocess_detections`: RoIHeadsSNN.postprocess_detections (roi_heads.py:1075-1176) with its
    per-detection Python loop (`for i in inds...: torch.where(inds_bg == i)`, roi_heads.py:1143-1146; one device
    sync per kept detection) replaced by a scatter on a mask.  Same outputs, same order.

Both are inference-only and keep the reference's (modified) return values: per-image pre-NMS proposals +
objectness, background boxes kept, `all_scores` / `all_boxes`.
"""
import ctypes
import math
import types
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor
from torchvision.ops import boxes as box_ops

from . import _lib


# --------------------------------------------------------------------------- rank 4
def postprocess_detections(class_logits: Tensor, box_regression: Tensor, proposals: List[Tensor],
                           image_shapes: List[Tuple[int, int]], box_coder, score_thresh: float, nms_thresh: float,
                           detections_per_img: int, use_kernel=None):
    """Vectorised RoIHeadsSNN.postprocess_detections (roi_heads.py:1075-1176); returns the same 5 lists:
    boxes, scores, labels (objects first, then every surviving background box), all_scores, all_boxes.
    On CUDA tensors everything after the softmax and the box decode is one kernel launch and one read of the
    per-image counts (`snn_det_postprocess`; `use_kernel=False` keeps the torch/torchvision ops below, which the
    kernel reproduces bit for bit -- tests/test_detection_post.py)."""
    device = class_logits.device
    num_classes = class_logits.shape[-1]
    boxes_per_image = [b.shape[0] for b in proposals]
    pred_boxes = box_coder.decode(box_regression, proposals)
    pred_scores = F.softmax(class_logits, -1)
    if use_kernel is None:
        use_kernel = _det_kernel_eligible(pred_scores, pred_boxes, boxes_per_image, num_classes)
    if use_kernel:
        return _postprocess_detections_kernel(pred_scores, pred_boxes, boxes_per_image, image_shapes, score_thresh,
                                              nms_thresh, detections_per_img)
    pred_boxes_list = pred_boxes.split(boxes_per_image, 0)
    pred_scores_list = pred_scores.split(boxes_per_image, 0)

    all_boxes, all_scores, all_labels, all_scores_all_classes, all_pre_nms_boxes = [], [], [], [], []
    for boxes, scores, image_shape in zip(pred_boxes_list, pred_scores_list, image_shapes):
        boxes = box_ops.clip_boxes_to_image(boxes, image_shape)
        labels = torch.arange(num_classes, device=device).view(1, -1).expand_as(scores)
        boxes_all_classes = boxes.detach().clone()
        scores_all_classes = scores.detach().clone()

        boxes_bg = boxes[:, 0].detach().reshape(-1, 4)
        scores_bg = scores[:, 0].detach().reshape(-1)
        labels_bg = labels[:, 0].reshape(-1)
        boxes = boxes[:, 1:].reshape(-1, 4)
        scores = scores[:, 1:].reshape(-1)
        labels = labels[:, 1:].reshape(-1)

        inds = torch.where(scores > score_thresh)[0]
        boxes, scores, labels = boxes[inds], scores[inds], labels[inds]

        # A background box is dropped when any class of the same RoI passed the threshold.  The reference loops over
        # `inds` in Python (roi_heads.py:1143-1146); the same mask is one scatter.  (`scores_bg >= 0` keeps NaN rows
        # out exactly as the reference's torch.where does.)
        roi_of_kept = torch.div(inds, num_classes - 1, rounding_mode="trunc")
        keep_bg_mask = scores_bg >= 0
        keep_bg_mask[roi_of_kept] = False
        inds_bg = torch.where(keep_bg_mask)[0]
        boxes_bg, scores_bg, labels_bg = boxes_bg[inds_bg], scores_bg[inds_bg], labels_bg[inds_bg]

        keep = box_ops.remove_small_boxes(boxes, min_size=1e-2)
        boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
        keep_bg = box_ops.remove_small_boxes(boxes_bg, min_size=1e-2)
        boxes_bg, scores_bg, labels_bg = boxes_bg[keep_bg], scores_bg[keep_bg], labels_bg[keep_bg]

        keep = box_ops.batched_nms(boxes, scores, labels, nms_thresh)
        keep_bg = box_ops.batched_nms(boxes_bg, scores_bg, labels_bg, nms_thresh)
        keep = keep[:detections_per_img]
        boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
        boxes_bg, scores_bg, labels_bg = boxes_bg[keep_bg], scores_bg[keep_bg], labels_bg[keep_bg]

        all_boxes.append(torch.cat((boxes, boxes_bg), dim=0))
        all_scores.append(torch.cat((scores, scores_bg), dim=0))
        all_labels.append(torch.cat((labels, labels_bg), dim=0))
        all_scores_all_classes.append(scores_all_classes)
        all_pre_nms_boxes.append(boxes_all_classes)
    return all_boxes, all_scores, all_labels, all_scores_all_classes, all_pre_nms_boxes


def _det_kernel_eligible(pred_scores, pred_boxes, boxes_per_image, num_classes) -> bool:
    if not (pred_scores.is_cuda and pred_boxes.is_cuda and pred_scores.dtype == torch.float32
            and pred_boxes.dtype == torch.float32 and num_classes >= 2 and len(boxes_per_image) >= 1):
        return False
    return max(boxes_per_image) * (num_classes - 1) <= _lib.load().snn_det_postprocess_max_candidates()


def _postprocess_detections_kernel(pred_scores, pred_boxes, boxes_per_image, image_shapes, score_thresh, nms_thresh,
                                   detections_per_img, objects_only=False):
    lib = _lib.load()
    dev = pred_scores.device
    n_img, C = len(boxes_per_image), pred_scores.shape[-1]
    pred_scores = pred_scores.contiguous()
    pred_boxes = pred_boxes.reshape(-1, C, 4).contiguous()
    cap = int(detections_per_img) + max(boxes_per_image)
    all_boxes = torch.empty_like(pred_boxes)
    out_boxes = torch.empty((n_img, cap, 4), dtype=torch.float32, device=dev)
    out_scores = torch.empty((n_img, cap), dtype=torch.float32, device=dev)
    out_labels = torch.empty((n_img, cap), dtype=torch.int64, device=dev)
    counts = torch.empty((n_img, 2), dtype=torch.int32, device=dev)
    IntArr = ctypes.c_int * n_img
    with torch.cuda.device(dev):
        rc = lib.snn_det_postprocess(pred_scores.data_ptr(), pred_boxes.data_ptr(), IntArr(*boxes_per_image),
                                     IntArr(*[int(s[0]) for s in image_shapes]), IntArr(*[int(s[1]) for s in image_shapes]),
                                     n_img, C, float(score_thresh), float(nms_thresh), 1e-2, int(detections_per_img), cap,
                                     all_boxes.data_ptr(), out_boxes.data_ptr(), out_scores.data_ptr(),
                                     out_labels.data_ptr(), counts.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "snn_det_postprocess")
    kept = (counts[:, 0] if objects_only else counts.sum(dim=1)).tolist()      # the one synchronisation of the call
    boxes = [out_boxes[b, :kept[b]] for b in range(n_img)]
    scores = [out_scores[b, :kept[b]] for b in range(n_img)]
    labels = [out_labels[b, :kept[b]] for b in range(n_img)]
    return (boxes, scores, labels, list(pred_scores.split(boxes_per_image, 0)), list(all_boxes.split(boxes_per_image, 0)))


def patch_postprocess(roi_heads):
    """Bind the vectorised post-processing to a RoIHeadsSNN instance (same signature as the reference method)."""
    def _pp(self, class_logits, box_regression, proposals, image_shapes):
        return postprocess_detections(class_logits, box_regression, proposals, image_shapes, self.box_coder,
                                      self.score_thresh, self.nms_thresh, self.detections_per_img)
    roi_heads.postprocess_detections = types.MethodType(_pp, roi_heads)
    return roi_heads


def patch_postprocess_torchvision(roi_heads):
    """The same for a stock torchvision RoIHeads, whose postprocess_detections returns (boxes, scores, labels) without
    the reference's background boxes and `all_*` lists -- the objects part of the reference's function, which is where
    the reference took it from (roi_heads.py:1075-1134 = torchvision's).  CUDA tensors within the kernel's limits run
    `snn_det_postprocess` and return its object rows; anything else goes to torchvision's own method."""
    original = roi_heads.postprocess_detections

    def _pp(self, class_logits, box_regression, proposals, image_shapes):
        boxes_per_image = [b.shape[0] for b in proposals]
        if not (class_logits.is_cuda and class_logits.dtype == torch.float32 and len(boxes_per_image) >= 1
                and class_logits.shape[-1] >= 2 and max(boxes_per_image) * (class_logits.shape[-1] - 1)
                <= _lib.load().snn_det_postprocess_max_candidates()):
            return original(class_logits, box_regression, proposals, image_shapes)
        pred_boxes = self.box_coder.decode(box_regression, proposals)
        pred_scores = F.softmax(class_logits, -1)
        boxes, scores, labels, _, _ = _postprocess_detections_kernel(pred_scores, pred_boxes, boxes_per_image, image_shapes,
                                                                    self.score_thresh, self.nms_thresh,
                                                                    self.detections_per_img, objects_only=True)
        return boxes, scores, labels
    roi_heads.postprocess_detections = types.MethodType(_pp, roi_heads)
    return roi_heads


# --------------------------------------------------------------------------- rank 1
def nchw_index_to_reference_order(idx: Tensor, A: int, H: int, W: int) -> Tensor:
    """Position inside a level's [A][H][W] logits -> position in the reference's (H, W, A) flattening
    (permute_and_flatten, rpn.py:248-259)."""
    a = torch.div(idx, H * W, rounding_mode="floor")
    rem = idx - a * (H * W)
    return rem * A + a


TOPK_KERNEL_MAX_K = 2048          # csrc/aux_kernels.cuh kTopkMaxK
_TOPK_WS = {}                     # (device, levels, images) -> workspace of the top-k kernels


def rpn_select_proposals(objectness: Sequence[Tensor], pred_bbox_deltas: Sequence[Tensor], cell_anchors: Sequence[Tensor],
                         strides: Sequence[Tuple[int, int]], pre_nms_top_n: int):
    """Per-level top-k on the head's NCHW logits + decode of the selected anchors only (CUDA).

    objectness[l] [N,A,H,W], pred_bbox_deltas[l] [N,4A,H,W] (CUDA fp32, as RPNHeadSNN returns them),
    cell_anchors[l] [A,4], strides[l] = (stride_h, stride_w).
    Returns proposals [N,K,4], objectness_prob [N,K], levels [N,K] (int64) and ref_index [N,K] (the anchor index in
    the reference's concatenated order), K = sum_l min(pre_nms_top_n, A*H*W), level-major like the reference."""
    lib = _lib.load()
    L = len(objectness)
    if L == 0:
        raise RuntimeError("rpn_select_proposals: no feature levels")
    dev = objectness[0].device
    if not objectness[0].is_cuda:
        raise RuntimeError("rpn_select_proposals: expected CUDA tensors (B200); there is no CPU fallback")
    N, A = objectness[0].shape[:2]
    logits = [o.detach().float().contiguous() for o in objectness]
    deltas = [d.detach().float().contiguous() for d in pred_bbox_deltas]
    bases = [c.detach().to(device=dev, dtype=torch.float32).contiguous() for c in cell_anchors]
    VP = ctypes.c_void_p * L
    IA = ctypes.c_int * L
    # top-k on unique integer keys (logit, then lowest index in the reference's (H, W, A) order): the same selection and
    # order as the reference's top-k on its permuted tensor, ties included -- pixels where no shared_lif neuron spiked
    # have exactly-zero membranes for every anchor, so tie groups at the k boundary are real
    sizes = [o[0].numel() for o in logits]
    ks = [min(int(pre_nms_top_n), n) for n in sizes]
    lvls = [torch.full((k,), l, dtype=torch.int64, device=dev) for l, k in enumerate(ks)]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    Hs, Ws = IA(*[t.shape[2] for t in logits]), IA(*[t.shape[3] for t in logits])
    if int(pre_nms_top_n) <= TOPK_KERNEL_MAX_K:
        # all levels and images at once: radix select + sort in the library (eight launches)
        idx = torch.empty(N, sum(ks), dtype=torch.int64, device=dev)
        ws = _TOPK_WS.get((str(dev), L, N))
        if ws is None:
            ws = _TOPK_WS[(str(dev), L, N)] = torch.empty(lib.snn_rpn_topk_workspace_bytes(L, N), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.snn_rpn_topk_select(VP(*[t.data_ptr() for t in logits]), Hs, Ws, L, N, A, int(pre_nms_top_n),
                                         ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), st)
        _lib.check(rc, "snn_rpn_topk_select")
    else:
        # larger k: the keys as a tensor + torch.topk per level
        keys_flat = torch.empty(N * sum(sizes), dtype=torch.int64, device=dev)
        keys, off = [], 0
        for n_anchors in sizes:
            keys.append(keys_flat[off:off + N * n_anchors].view(N, n_anchors)); off += N * n_anchors
        with torch.cuda.device(dev):
            rc = lib.snn_rpn_topk_keys(VP(*[t.data_ptr() for t in logits]), Hs, Ws, L, N, A, VP(*[t.data_ptr() for t in keys]), st)
        _lib.check(rc, "snn_rpn_topk_keys")
        idx = torch.cat([kt.topk(k, dim=1)[1] for kt, k in zip(keys, ks)], dim=1).contiguous()
    K = idx.shape[1]
    boxes = torch.empty(N, K, 4, device=dev, dtype=torch.float32)
    scores = torch.empty(N, K, device=dev, dtype=torch.float32)
    ref_index = torch.empty(N, K, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        rc = lib.snn_rpn_decode_selected(
            VP(*[t.data_ptr() for t in logits]), VP(*[t.data_ptr() for t in deltas]), VP(*[t.data_ptr() for t in bases]),
            IA(*[t.shape[2] for t in logits]), IA(*[t.shape[3] for t in logits]),
            IA(*[int(s[0]) for s in strides]), IA(*[int(s[1]) for s in strides]), IA(*ks), L, N, A,
            ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(boxes.data_ptr()), ctypes.c_void_p(scores.data_ptr()),
            None, ctypes.c_void_p(ref_index.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "snn_rpn_decode_selected")
    levels = torch.cat(lvls).reshape(1, -1).expand(N, -1)
    return boxes, scores, levels, ref_index


_RPN_NMS_WS = {}                  # (device, levels, images) -> workspace of the proposal-filter kernel
RPN_NMS_MAX_LEVEL = 2048          # entries per (image, level) snn_rpn_nms sorts in shared memory
RPN_NMS_MAX_KEPT = 8192           # keepers per image its merging block sorts


def filter_selected(proposals: Tensor, objectness_prob: Tensor, levels: Tensor, image_shapes: List[Tuple[int, int]],
                    min_size: float, score_thresh: float, nms_thresh: float, post_nms_top_n: int,
                    level_sizes: Sequence[int] = None, use_kernel=None):
    """The tail of RegionProposalNetwork.filter_proposals (rpn.py:493-525) on the already selected / decoded entries.
    With `level_sizes` (entries per level and image of the level-major lists, as rpn_select_proposals lays them out) and
    CUDA tensors it is one launch of `snn_rpn_nms` and one read of the per-image counts; `use_kernel=False` keeps the
    torchvision ops below, whose result the kernel reproduces bit for bit (tests/test_detection_post.py)."""
    pre_nms = [{"proposals": prop, "objectness": objectness_prob[i]} for i, prop in enumerate(proposals)]
    if use_kernel is None:
        use_kernel = (level_sizes is not None and proposals.is_cuda and proposals.dtype == torch.float32
                      and objectness_prob.dtype == torch.float32 and proposals.shape[0] >= 1
                      and max(level_sizes) <= RPN_NMS_MAX_LEVEL
                      and sum(min(int(post_nms_top_n), int(k)) for k in level_sizes) <= RPN_NMS_MAX_KEPT)
    if use_kernel:
        if level_sizes is None or sum(level_sizes) != proposals.shape[1]:
            raise RuntimeError("filter_selected: level_sizes must add up to the entries per image")
        lib = _lib.load()
        dev = proposals.device
        N, L, post_n = proposals.shape[0], len(level_sizes), int(post_nms_top_n)
        props = proposals.contiguous()
        probs = objectness_prob.contiguous()
        out_boxes = torch.empty((N, post_n, 4), dtype=torch.float32, device=dev)
        out_scores = torch.empty((N, post_n), dtype=torch.float32, device=dev)
        counts = torch.empty((N,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws = _RPN_NMS_WS.get((str(dev), L, N))
            if ws is None:
                ws = _RPN_NMS_WS[(str(dev), L, N)] = torch.empty(lib.snn_rpn_nms_workspace_bytes(L, N), dtype=torch.uint8, device=dev)
            IntL, IntN = ctypes.c_int * L, ctypes.c_int * N
            rc = lib.snn_rpn_nms(props.data_ptr(), probs.data_ptr(), IntL(*[int(k) for k in level_sizes]),
                                 IntN(*[int(s[0]) for s in image_shapes]), IntN(*[int(s[1]) for s in image_shapes]),
                                 L, N, float(min_size), float(score_thresh), float(nms_thresh), post_n,
                                 out_boxes.data_ptr(), out_scores.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "snn_rpn_nms")
        kept = counts.tolist()                             # the one synchronisation of the call
        return [out_boxes[i, :kept[i]] for i in range(N)], [out_scores[i, :kept[i]] for i in range(N)], pre_nms
    final_boxes, final_scores = [], []
    for boxes, scores, lvl, img_shape in zip(proposals, objectness_prob, levels, image_shapes):
        boxes = box_ops.clip_boxes_to_image(boxes, img_shape)
        keep = box_ops.remove_small_boxes(boxes, min_size)
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = torch.where(scores >= score_thresh)[0]
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = box_ops.batched_nms(boxes, scores, lvl, nms_thresh)
        keep = keep[:post_nms_top_n]
        final_boxes.append(boxes[keep]); final_scores.append(scores[keep])
    return final_boxes, final_scores, pre_nms


def fast_rpn_forward(rpn, images, features: Dict[str, Tensor]):
    """Inference-time replacement of RegionProposalNetwork.forward (rpn.py:563-703): returns (boxes, extras) with
    extras = the per-image pre-NMS proposals + objectness, as the reference's modified forward does in eval mode."""
    feats = list(features.values())
    objectness, pred_bbox_deltas = rpn.head(feats)
    image_size = images.tensors.shape[-2:]
    strides = [(image_size[0] // f.shape[-2], image_size[1] // f.shape[-1]) for f in feats]
    props, probs, levels, _ = rpn_select_proposals(objectness, pred_bbox_deltas, rpn.anchor_generator.cell_anchors,
                                                   strides, rpn.pre_nms_top_n())
    level_sizes = [min(int(rpn.pre_nms_top_n()), int(o.shape[1] * o.shape[2] * o.shape[3])) for o in objectness]
    boxes, _scores, pre_nms = filter_selected(props, probs, levels, images.image_sizes, rpn.min_size, rpn.score_thresh,
                                              rpn.nms_thresh, rpn.post_nms_top_n(), level_sizes=level_sizes)
    return boxes, pre_nms


def attach_fast_postprocessing(model):
    """Patch a Faster R-CNN built like the reference's (model.py): eval-mode `rpn.forward` selects proposals from the
    head's native layout, and `roi_heads.postprocess_detections` is the vectorised one.  Training keeps the originals."""
    rpn = model.rpn
    original = rpn.forward
    reference_style = hasattr(model.roi_heads, "box_head_and_predictor")

    def _forward(self, images, features, targets=None):
        if self.training:
            return original(images, features, targets)
        boxes, pre_nms = fast_rpn_forward(self, images, features)
        # the reference's modified RPN returns the pre-NMS proposals in place of the (empty) loss dict;
        # a stock torchvision GeneralizedRCNN expects a dict there
        return boxes, (pre_nms if reference_style else {})

    rpn.forward = types.MethodType(_forward, rpn)
    if reference_style:                                              # the reference's RoIHeadsSNN
        patch_postprocess(model.roi_heads)
    else:
        patch_postprocess_torchvision(model.roi_heads)
    return model


BBOX_XFORM_CLIP = math.log(1000.0 / 16)


# --------------------------------------------------------------------------- rank 2
class FusedRoIAlignEncoder(torch.nn.Module):
    """Drop-in for `RoIHeadsSNN.box_roi_pool` (a torchvision MultiScaleRoIAlign; roi_heads.py:1217) when the box head
    is the B200 `FastRCNNPredictorSNNFull`: RoIAlign and the head's constant-current encoder run in ONE kernel
    (csrc/aux_kernels.cuh::roi_align_encode_kernel), so the pooled [R, C, 7, 7] fp32 tensor (50 MB per image written
    and read back) never exists.  forward(features, proposals, image_shapes) has the pooler's signature and returns
    `EncodedRoIs` (spike-train words [R, C*7*7] for num_steps - 1 encoder steps), which the head's forward accepts in
    place of the pooled tensor.  Level assignment and scales are torchvision's own (LevelMapper, _setup_scales)."""

    def __init__(self, featmap_names: Sequence[str], output_size: int, sampling_ratio: int, num_steps: int,
                 canonical_scale: int = 224, canonical_level: int = 4):
        super().__init__()
        self.featmap_names = list(featmap_names)
        self.output_size = int(output_size[0] if isinstance(output_size, (tuple, list)) else output_size)
        self.sampling_ratio = int(sampling_ratio)
        self.num_steps = int(num_steps)
        self.canonical_scale, self.canonical_level = canonical_scale, canonical_level
        self.return_pooled = False                      # tests: also keep the RoIAlign values
        self.last_pooled = None

    @classmethod
    def from_pooler(cls, pooler, num_steps: int):
        return cls(pooler.featmap_names, pooler.output_size, pooler.sampling_ratio, num_steps,
                   getattr(pooler, "canonical_scale", 224), getattr(pooler, "canonical_level", 4))

    @torch.no_grad()
    def forward(self, features: Dict[str, Tensor], proposals: List[Tensor], image_shapes: List[Tuple[int, int]]):
        from torchvision.ops.poolers import _setup_scales, _convert_to_roi_format
        from .heads import EncodedRoIs, _TRAIN_DTYPE
        lib = _lib.load()
        feats = [features[k].detach().float().contiguous() for k in self.featmap_names if k in features]
        if not feats or not feats[0].is_cuda:
            raise RuntimeError("FusedRoIAlignEncoder: expected CUDA feature maps (B200); there is no CPU fallback")
        dev = feats[0].device
        scales, mapper = _setup_scales(feats, image_shapes, self.canonical_scale, self.canonical_level)
        rois = _convert_to_roi_format(proposals).float().contiguous()              # [R, 5]
        R, C, P = rois.shape[0], feats[0].shape[1], self.output_size
        T_live = self.num_steps - 1
        wb = 1 if T_live <= 8 else 2 if T_live <= 16 else 4
        words = torch.empty(R, C * P * P, device=dev, dtype=_TRAIN_DTYPE[wb])
        pooled = torch.empty(R, C * P * P, device=dev, dtype=torch.float32) if self.return_pooled else None
        if R > 0:
            levels = (mapper(proposals) if len(feats) > 1 else torch.zeros(R, dtype=torch.int64, device=dev)).to(torch.int32).contiguous()
            L = len(feats)
            VP, IA, FA = ctypes.c_void_p * L, ctypes.c_int * L, ctypes.c_float * L
            with torch.cuda.device(dev):
                rc = lib.snn_roi_align_encode(
                    VP(*[f.data_ptr() for f in feats]), IA(*[f.shape[2] for f in feats]), IA(*[f.shape[3] for f in feats]),
                    FA(*[float(s) for s in scales]), L, C, ctypes.c_void_p(rois.data_ptr()), ctypes.c_void_p(levels.data_ptr()),
                    R, P, self.sampling_ratio, T_live, ctypes.c_void_p(words.data_ptr()),
                    ctypes.c_void_p(pooled.data_ptr()) if pooled is not None else None,
                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc, "snn_roi_align_encode")
        self.last_pooled = pooled
        return EncodedRoIs(words, self.num_steps)


def attach_fused_roi_pool(model):
    """Replace `roi_heads.box_roi_pool` by the fused RoIAlign + encoder (eval only; the reference's RoIHeadsSNN, whose
    forward hands the pooler's output straight to `box_head_and_predictor`, roi_heads.py:1217-1230)."""
    rh = model.roi_heads
    head = getattr(rh, "box_head_and_predictor", None)
    if head is None and isinstance(getattr(rh, "box_head", None), torch.nn.Identity):
        head = getattr(rh, "box_predictor", None)        # stock torchvision RoIHeads as attach_snn_heads leaves them
    if head is None or not hasattr(head, "num_steps"):
        raise RuntimeError("attach_fused_roi_pool: needs the reference's RoIHeadsSNN (box_head_and_predictor) or a "
                           "torchvision RoIHeads prepared by attach_snn_heads")
    rh.box_roi_pool = FusedRoIAlignEncoder.from_pooler(rh.box_roi_pool, int(head.num_steps))
    return model
