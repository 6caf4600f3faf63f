"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/post_*.npz for the "next" rows.

Runs the UNMODIFIED reference classes on CPU:
  * /root/reference/roi_heads.py::RoIHeadsSNN.postprocess_detections on random logits / regressions / proposals;
  * /root/reference/rpn.py::RegionProposalNetwork.forward (eval) with a stub head that returns fixed tensors.
Run in the build container only (needs /root/reference):  python oracle/gen_golden_post.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "norse_shim"))
sys.path.insert(0, "/root/reference")


def gen_postprocess():
    import roi_heads as ref_roi
    torch.manual_seed(11)
    C, per_img = 9, [150, 90]
    R = sum(per_img)
    rh = ref_roi.RoIHeadsSNN(None, None, 0.5, 0.5, 512, 0.25, None, score_thresh=0.4, nms_thresh=0.5, detections_per_img=100)
    logits = torch.randn(R, C) * 2.5
    logits[:, 0] += 1.0                                   # background often, but not always, the top class
    reg = torch.randn(R, 4 * C) * 0.5
    shapes = [(768, 1536), (700, 1400)]
    props = []
    for n, (h, w) in zip(per_img, shapes):
        xy = torch.rand(n, 2) * torch.tensor([w * 0.8, h * 0.8])
        wh = torch.rand(n, 2) * torch.tensor([w * 0.3, h * 0.3]) + 4.0
        props.append(torch.cat([xy, xy + wh], dim=1))
    with torch.no_grad():
        boxes, scores, labels, all_scores, all_boxes = rh.postprocess_detections(logits, reg, props, shapes)
    out = {"logits": logits.numpy(), "reg": reg.numpy(), "shapes": np.array(shapes), "per_img": np.array(per_img),
           "score_thresh": 0.4, "nms_thresh": 0.5, "detections_per_img": 100}
    for i in range(len(per_img)):
        out[f"props{i}"] = props[i].numpy()
        out[f"boxes{i}"] = boxes[i].numpy(); out[f"scores{i}"] = scores[i].numpy(); out[f"labels{i}"] = labels[i].numpy()
        out[f"all_scores{i}"] = all_scores[i].numpy(); out[f"all_boxes{i}"] = all_boxes[i].numpy()
    np.savez_compressed(os.path.join(OUT, "post_detections.npz"), **out)
    print("post_detections:", [b.shape[0] for b in boxes], "kept (objects + background)")


def gen_rpn():
    import rpn as ref_rpn
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    torch.manual_seed(12)
    N, A = 2, 3
    levels = [(24, 32), (12, 16), (6, 8)]
    img = (192, 256)
    sizes = ((32,), (64,), (128,))
    ag = AnchorGenerator(sizes, ((0.5, 1.0, 2.0),) * len(sizes))
    obj = [torch.randn(N, A, h, w) * 2 for (h, w) in levels]
    dlt = [torch.randn(N, 4 * A, h, w) * 0.4 for (h, w) in levels]

    class Stub(torch.nn.Module):
        def forward(self, feats):
            return obj, dlt

    net = ref_rpn.RegionProposalNetwork(ag, Stub(), 0.7, 0.3, 256, 0.5, dict(training=2000, testing=200),
                                        dict(training=2000, testing=100), 0.7, score_thresh=0.0)
    net.eval()
    images = ImageList(torch.zeros(N, 3, *img), [(192, 256), (180, 240)])
    feats = {str(i): torch.zeros(N, 8, h, w) for i, (h, w) in enumerate(levels)}
    with torch.no_grad():
        boxes, extras = net(images, feats)
    out = {"levels": np.array(levels), "img": np.array(img), "image_sizes": np.array([(192, 256), (180, 240)]),
           "pre_nms_top_n": 200, "post_nms_top_n": 100, "nms_thresh": 0.7, "min_size": float(net.min_size)}
    for l in range(len(levels)):
        out[f"obj{l}"] = obj[l].numpy(); out[f"dlt{l}"] = dlt[l].numpy()
        out[f"cell{l}"] = ag.cell_anchors[l].numpy()
    for i in range(N):
        out[f"boxes{i}"] = boxes[i].numpy()
        out[f"pre_props{i}"] = extras[i]["proposals"].numpy(); out[f"pre_obj{i}"] = extras[i]["objectness"].numpy()
    np.savez_compressed(os.path.join(OUT, "post_rpn.npz"), **out)
    print("post_rpn:", [b.shape[0] for b in boxes], "proposals after NMS;", extras[0]["proposals"].shape[0], "before")


if __name__ == "__main__":
    gen_postprocess()
    gen_rpn()
