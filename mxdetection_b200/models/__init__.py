"""mxdetection/models (/root/reference/README.md:26-33): only the two box-arithmetic modules on the hot path."""
from . import roi_extractors, rpn_heads  # noqa: F401
