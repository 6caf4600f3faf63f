#!/usr/bin/env python
"""Whole-model context for the head metric ("an end-to-end GeneralizedRCNN number may be reported beside it"):
a stock torchvision Faster R-CNN ResNet50-FPN (the architecture model.py builds; random init, no download) with the B200
spiking heads attached, on 2 Cityscapes-sized images (1024 x 2048 -> 768 x 1536), T 8/12, 9 classes, fp16x2 weights.
Times one eval forward (CUDA events, ms per batch of 2) with (a) only the two heads swapped, (b) + the fast proposal
path (attach_fast_postprocessing: selection from the NCHW logits, snn_rpn_nms), (c) + RoIAlign and encoder fused
(attach_fused_roi_pool), and per stage with forward hooks for (c).  A random-init frozen-BN backbone yields features with
std ~ 50 (a stress case for the encoder's firing rate), which does not change the heads' dense tensor work.
Usage: python profiles/bench_full_model.py [--out file.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snn_automotive_object_detection_b200 as S  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from torchvision.models.detection import fasterrcnn_resnet50_fpn
    torch.manual_seed(0)
    model = fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, min_size=768, max_size=1536)
    S.attach_snn_heads(model, num_steps_rpn=8, num_steps_detector=12, num_classes=9, mode="fp16x2")
    model = model.cuda().eval()
    imgs = [torch.rand(3, 1024, 2048, device="cuda") for _ in range(2)]

    def run():
        with torch.no_grad():
            return model(imgs)

    out = {"config": "torchvision fasterrcnn_resnet50_fpn (random init) + B200 SNN heads, 2 x 1024x2048 -> 768x1536, T 8/12, 9 classes, fp16x2"}
    out["a_heads_only_ms"] = timed(run)
    S.attach_fast_postprocessing(model)
    out["b_plus_fast_proposals_ms"] = timed(run)
    S.attach_fused_roi_pool(model)
    out["c_plus_fused_roi_pool_ms"] = timed(run)
    out["detections"] = [int(d["boxes"].shape[0]) for d in run()]

    # per stage (configuration c): CUDA events around the sub-modules
    stages = {"transform": model.transform, "backbone": model.backbone, "rpn": model.rpn, "rpn_head": model.rpn.head,
              "roi_heads": model.roi_heads, "box_roi_pool": model.roi_heads.box_roi_pool,
              "box_predictor": model.roi_heads.box_predictor}
    ev = {k: [] for k in stages}
    hooks = []
    for name, mod in stages.items():
        def pre(m, args, name=name):
            e = torch.cuda.Event(enable_timing=True); e.record(); ev[name].append([e, None])
        def post(m, args, outp, name=name):
            e = torch.cuda.Event(enable_timing=True); e.record(); ev[name][-1][1] = e
        hooks.append(mod.register_forward_pre_hook(pre)); hooks.append(mod.register_forward_hook(post))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    for h in hooks:
        h.remove()
    st = {k: sum(a_.elapsed_time(b_) for a_, b_ in v) / len(v) for k, v in ev.items() if v}
    st["rpn_without_head"] = st["rpn"] - st["rpn_head"]
    st["roi_heads_without_pool_and_predictor"] = st["roi_heads"] - st["box_roi_pool"] - st["box_predictor"]
    out["stage_ms_config_c"] = {k: round(v, 3) for k, v in st.items()}
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
