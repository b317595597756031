"""mxdetection/core/mask (/root/reference/README.md:18): mask targets."""
from .mask_target import mask_target  # noqa: F401
