"""mxdetection/core/mask (/root/reference/README.md:18): mask targets and mask paste."""
from .mask_target import mask_target, paste_masks  # noqa: F401
