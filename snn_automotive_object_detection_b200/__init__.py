"""B200-native spiking detection heads (drop-in for the reference's Norse path)."""
from .heads import RPNHeadSNN, FastRCNNPredictorSNNFull, EncoderParameters, unpack_trains  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["RPNHeadSNN", "FastRCNNPredictorSNNFull", "EncoderParameters", "unpack_trains"]
