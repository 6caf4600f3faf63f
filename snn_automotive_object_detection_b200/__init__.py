"""B200-native spiking detection heads (drop-in for the reference's Norse path)."""
from .heads import RPNHeadSNN, FastRCNNPredictorSNNFull, EncoderParameters, EncodedRoIs, unpack_trains  # noqa: F401
from .plugin import attach_snn_heads  # noqa: F401
from .rates import rpn_spike_rates_and_flops, box_spike_rates_and_flops, energy_ratio, energy_report  # noqa: F401
from .detection_post import (postprocess_detections, patch_postprocess, rpn_select_proposals, filter_selected,  # noqa: F401
                             fast_rpn_forward, attach_fast_postprocessing, FusedRoIAlignEncoder, attach_fused_roi_pool)
from . import _lib, parallel, detection_post  # noqa: F401

__all__ = ["RPNHeadSNN", "FastRCNNPredictorSNNFull", "EncoderParameters", "unpack_trains", "attach_snn_heads",
           "rpn_spike_rates_and_flops", "box_spike_rates_and_flops", "energy_ratio", "energy_report", "parallel",
           "postprocess_detections", "patch_postprocess", "rpn_select_proposals", "filter_selected", "fast_rpn_forward",
           "attach_fast_postprocessing", "FusedRoIAlignEncoder", "attach_fused_roi_pool", "EncodedRoIs"]
