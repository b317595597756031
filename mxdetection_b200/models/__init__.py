"""mxdetection/models (/root/reference/README.md:26-33): the box-arithmetic modules on and next to the hot path."""
from . import roi_extractors, rpn_heads, bbox_heads, mask_heads  # noqa: F401
