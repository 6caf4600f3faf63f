"""norse.torch namespace (oracle shim) -- names imported at rpn.py:16-19 and
faster_rcnn.py:24-27 of the reference."""
from .functional.lif import LIFParameters, LIFFeedForwardState, lif_feed_forward_step, lif_current_encoder
from .functional.leaky_integrator import LIParameters, LIState, li_feed_forward_step
from .module.lif import LIFCell
from .module.leaky_integrator import LICell

__all__ = [
    "LIFParameters", "LIFFeedForwardState", "lif_feed_forward_step", "lif_current_encoder",
    "LIParameters", "LIState", "li_feed_forward_step", "LIFCell", "LICell",
]
