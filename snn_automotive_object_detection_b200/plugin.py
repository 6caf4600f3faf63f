"""Plug the B200 spiking heads into a Faster R-CNN, exactly where the reference puts its Norse heads.

The reference builds a torchvision-style FasterRCNN and swaps the modules by assignment
(model.py:61-68, 127-130, 186-187): `rpn.head = RPNHeadSNN(...)` and
`roi_heads.box_head_and_predictor = FastRCNNPredictorSNNFull(...)`.  attach_snn_heads() does the
same on either the reference's own classes (RoIHeadsSNN has `box_head_and_predictor`) or a stock
torchvision model (RoIHeads calls `box_predictor(box_head(x))`, so box_head becomes the identity).
GeneralizedRCNN.forward(images, targets=None) and everything around the heads stay untouched.
"""
from torch import nn

from .heads import RPNHeadSNN, FastRCNNPredictorSNNFull


def attach_snn_heads(model, num_steps_rpn: int, num_steps_detector: int, num_classes: int, rpn_snn: bool = True,
                     detector_snn: bool = True, only_one_bbox: bool = False, representation_size: int = 1024,
                     mode="fp32_exact"):
    out_channels = model.backbone.out_channels
    if rpn_snn:
        num_anchors = model.rpn.anchor_generator.num_anchors_per_location()[0]
        model.rpn.head = RPNHeadSNN(out_channels, num_anchors, num_steps_rpn, mode=mode)          # model.py:61-68
    if detector_snn:
        resolution = model.roi_heads.box_roi_pool.output_size[0]
        head = FastRCNNPredictorSNNFull(out_channels * resolution ** 2, representation_size, num_classes,
                                        num_steps_detector, only_one_bbox=only_one_bbox, mode=mode)   # model.py:127-130
        if hasattr(model.roi_heads, "box_head_and_predictor"):       # the reference's RoIHeadsSNN (roi_heads.py:1230)
            model.roi_heads.box_head_and_predictor = head
        else:                                                        # stock torchvision RoIHeads
            model.roi_heads.box_head = nn.Identity()
            model.roi_heads.box_predictor = head
    return model
