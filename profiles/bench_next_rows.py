#!/usr/bin/env python
"""Timing of the "next" rows at the Cityscapes batch-2 shapes, CUDA events, one JSON line.

  rpn_select  : per-level top-k on the head's NCHW outputs + decode of the selected anchors (ours) against the path
                the reference runs after RPNHeadSNN (rpn.py:636-670): permute/reshape of all logits and deltas,
                AnchorGenerator, BoxCoder.decode of every anchor, per-level top-k, gather, sigmoid -- restated here with
                the torchvision pieces the reference itself calls.
  postprocess : vectorised RoIHeadsSNN.postprocess_detections against the reference's per-detection Python loop
                (roi_heads.py:1143-1146, restated for timing only).
"""
import json
import os
import sys

import torch
from torchvision.models.detection import _utils as det_utils
from torchvision.models.detection.anchor_utils import AnchorGenerator
from torchvision.models.detection.image_list import ImageList

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snn_automotive_object_detection_b200 import detection_post as DP  # noqa: E402

LEVELS = [(192, 384), (96, 192), (48, 96), (24, 48), (12, 24)]
N, A, IMG = 2, 3, (768, 1536)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda")
    obj = [torch.randn(N, A, h, w, device=dev) for (h, w) in LEVELS]
    dlt = [torch.randn(N, 4 * A, h, w, device=dev) * 0.3 for (h, w) in LEVELS]
    ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)
    images = ImageList(torch.zeros(N, 3, *IMG, device=dev), [IMG] * N)
    feats = [torch.zeros(N, 1, h, w, device=dev) for (h, w) in LEVELS]
    coder = det_utils.BoxCoder((1.0, 1.0, 1.0, 1.0))
    strides = [(IMG[0] // h, IMG[1] // w) for (h, w) in LEVELS]

    def reference_path():
        anchors = ag(images, feats)
        o = torch.cat([x.view(N, -1, 1, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(N, -1, 1) for x in obj], dim=1).flatten(0, -2)
        d = torch.cat([x.view(N, -1, 4, *x.shape[-2:]).permute(0, 3, 4, 1, 2).reshape(N, -1, 4) for x in dlt], dim=1).reshape(-1, 4)
        props = coder.decode(d, anchors).view(N, -1, 4)
        o = o.reshape(N, -1)
        r, off = [], 0
        for ob in o.split([A * h * w for (h, w) in LEVELS], 1):
            _, idx = ob.topk(min(1000, ob.shape[1]), dim=1)
            r.append(idx + off); off += ob.shape[1]
        top = torch.cat(r, dim=1)
        b = torch.arange(N, device=dev)[:, None]
        return props[b, top], torch.sigmoid(o[b, top]), props, o

    def ours():
        return DP.rpn_select_proposals(obj, dlt, ag.cell_anchors, strides, 1000)

    # same entries?  compare through the reference-order anchor index each selected entry reports (equal logits at
    # two anchors may legitimately swap places between the two top-k calls)
    _, sr, all_props, all_logits = reference_path()
    po, so, _, ref_index = ours()
    b = torch.arange(N, device=dev)[:, None]
    pr = all_props[b, ref_index]
    assert torch.equal(torch.sigmoid(all_logits[b, ref_index]), so) or (torch.sigmoid(all_logits[b, ref_index]) - so).abs().max() <= 1e-6
    ds, db = (sr - so).abs().max().item(), ((pr - po).abs() / (1.0 + pr.abs())).max().item()
    assert ds <= 1e-6 and db <= 1e-5, f"fast path differs from the reference path: scores {ds}, boxes (relative) {db}"

    # ---- the tail of filter_proposals (rpn.py:493-525) on the selected entries: clip, small, threshold, NMS 0.7, top 1000
    po_, so_, lv_, _ = ours()

    def rpn_filter():
        return DP.filter_selected(po_, so_, lv_, [IMG] * N, 1e-3, 0.0, 0.7, 1000)

    LEVEL_SIZES = [min(1000, A * h * w) for (h, w) in LEVELS]

    def rpn_filter_kernel():
        return DP.filter_selected(po_, so_, lv_, [IMG] * N, 1e-3, 0.0, 0.7, 1000, level_sizes=LEVEL_SIZES)

    def rpn_filter():
        return DP.filter_selected(po_, so_, lv_, [IMG] * N, 1e-3, 0.0, 0.7, 1000, use_kernel=False)

    fa, fb = rpn_filter(), rpn_filter_kernel()
    assert all(torch.equal(x, y) for k in range(2) for x, y in zip(fa[k], fb[k])), "rpn filter kernel differs from the torch ops"
    t_rpn_filter, t_rpn_filter_kernel = timed(rpn_filter), timed(rpn_filter_kernel)
    n_after_nms = [int(b.shape[0]) for b in fb[0]]

    # ---- detection post-processing: 2 x 1000 RoIs, 9 classes, thresholds of model.py:98-99
    C, R = 9, 1000
    logits = torch.randn(N * R, C, device=dev) * 2.5
    reg = torch.randn(N * R, 4 * C, device=dev) * 0.5
    xy = torch.rand(N * R, 2, device=dev) * torch.tensor([IMG[1] * 0.8, IMG[0] * 0.8], device=dev)
    wh = torch.rand(N * R, 2, device=dev) * 300 + 4
    props = list(torch.cat([xy, xy + wh], dim=1).split(R))
    coder2 = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))

    def vectorised():
        return DP.postprocess_detections(logits, reg, props, [IMG] * N, coder2, 0.4, 0.5, 100, use_kernel=False)

    def kernel_path():         # softmax + decode in torch, everything after in one launch (snn_det_postprocess)
        return DP.postprocess_detections(logits, reg, props, [IMG] * N, coder2, 0.4, 0.5, 100)

    def softmax_and_decode_only():
        return coder2.decode(reg, props), torch.softmax(logits, -1)

    import ctypes
    from snn_automotive_object_detection_b200 import _lib
    lib_ = _lib.load()
    sc_, bx_ = torch.softmax(logits, -1).contiguous(), coder2.decode(reg, props).reshape(-1, C, 4).contiguous()
    cap_ = 100 + R
    bufs_ = [torch.empty_like(bx_), torch.empty(N, cap_, 4, device=dev), torch.empty(N, cap_, device=dev),
             torch.empty(N, cap_, dtype=torch.int64, device=dev), torch.empty(N, 2, dtype=torch.int32, device=dev)]
    IntArr = ctypes.c_int * N

    def det_kernel_only():
        rc = lib_.snn_det_postprocess(sc_.data_ptr(), bx_.data_ptr(), IntArr(*([R] * N)), IntArr(*([IMG[0]] * N)),
                                     IntArr(*([IMG[1]] * N)), N, C, 0.4, 0.5, 1e-2, 100, cap_, *[t.data_ptr() for t in bufs_],
                                     torch.cuda.current_stream().cuda_stream)
        assert rc == 0

    va, ka = vectorised(), kernel_path()
    assert all(torch.equal(x, y) for k in range(5) for x, y in zip(va[k], ka[k])), "kernel path differs from the torch ops"

    def loop_mask_only():      # the part the reference does per detection in Python (roi_heads.py:1136-1147)
        scores = torch.softmax(logits, -1)
        for sc in scores.split(R):
            s = sc[:, 1:].reshape(-1)
            inds = torch.where(s > 0.4)[0]
            which = torch.div(inds, C - 1, rounding_mode="trunc")
            inds_bg = torch.where(sc[:, 0] >= 0)[0]
            mask = torch.ones(*inds_bg.shape, dtype=torch.int32, device=dev)
            for i in which:
                pos = torch.where(inds_bg == i)[0]
                if len(pos):
                    mask[pos] = 0

    # ---- RoIAlign -> encoder: torchvision MultiScaleRoIAlign + the box head's encoder kernel, against the fused kernel
    import ctypes
    from collections import OrderedDict
    from torchvision.ops import MultiScaleRoIAlign
    from snn_automotive_object_detection_b200 import _lib
    fmaps = OrderedDict((str(i), torch.randn(N, 256, h, w, device=dev)) for i, (h, w) in enumerate(LEVELS[:4]))
    pooler = MultiScaleRoIAlign(featmap_names=["0", "1", "2", "3"], output_size=7, sampling_ratio=2)
    fused = DP.FusedRoIAlignEncoder.from_pooler(pooler, 12)
    lib = _lib.load()
    zbuf = torch.empty(N * R, 12544, dtype=torch.int16, device=dev)

    def pool_then_encode():
        x = pooler(fmaps, props, [IMG] * N).flatten(1).contiguous()
        lib.snn_encode_rows(ctypes.c_void_p(x.data_ptr()), N * R, 12544, 11, ctypes.c_void_p(zbuf.data_ptr()),
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

    def fused_pool():
        return fused(fmaps, props, [IMG] * N)

    # the fused kernel alone (C ABI call, inputs prepared): two against four channel planes in flight per thread
    from torchvision.ops.poolers import _setup_scales, _convert_to_roi_format
    fl = [fmaps[k] for k in ("0", "1", "2", "3")]
    scales, mapper = _setup_scales(fl, [IMG] * N, 224, 4)
    rois5 = _convert_to_roi_format(props).float().contiguous()
    lv = mapper(props).to(torch.int32).contiguous()
    words = torch.empty(N * R, 12544, dtype=torch.int16, device=dev)
    VP, IA, FA = ctypes.c_void_p * 4, ctypes.c_int * 4, ctypes.c_float * 4

    def roi_kernel_only():
        rc = lib.snn_roi_align_encode(VP(*[f.data_ptr() for f in fl]), IA(*[f.shape[2] for f in fl]), IA(*[f.shape[3] for f in fl]),
                                      FA(*[float(s_) for s_ in scales]), 4, 256, ctypes.c_void_p(rois5.data_ptr()),
                                      ctypes.c_void_p(lv.data_ptr()), N * R, 7, 2, 11, ctypes.c_void_p(words.data_ptr()), None,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0

    t_staged = timed(roi_kernel_only, iters=50)
    w_staged = words.clone()
    lib.snn_set_roi_kernel(1)
    t_thread = timed(roi_kernel_only, iters=50)
    lib.snn_set_roi_kernel(0)
    same_words = (w_staged == words).float().mean().item()
    line = {"rows": "SURVEY 8f-1 / 8f-2 / 8f-4",
            "roi_align_encode_ms": {"torchvision_roi_align_then_encoder": timed(pool_then_encode), "fused": timed(fused_pool),
                                    "kernel_only_unroll2": t_staged, "kernel_only_unroll4": t_thread,
                                    "words_equal_between_kernels": same_words},
            "shapes": "cityscapes batch 2 (294 624 anchors/img, 1000 RoIs/img)",
            "rpn_select_ms": {"reference_path": timed(reference_path), "ours": timed(ours)},
            "rpn_filter_tail_ms": {"torch_ops": t_rpn_filter, "ours": t_rpn_filter_kernel, "proposals_after_nms": n_after_nms},
            "postprocess_ms": {"reference_python_mask_loop_only": timed(loop_mask_only, iters=3), "torch_ops_whole_function": timed(vectorised),
                               "ours_whole_function": timed(kernel_path), "of_which_softmax_and_decode": timed(softmax_and_decode_only),
                               "of_which_the_kernel": timed(det_kernel_only, iters=50),
                               "detections_kept": [int(x.shape[0]) for x in ka[0]]}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
