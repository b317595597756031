"""mxdetection/core/anchor (/root/reference/README.md:16)."""
from .anchor_generator import AnchorGenerator, generate_anchors_mx  # noqa: F401
from .anchor_target import anchor_inside_flags, anchor_assign  # noqa: F401
