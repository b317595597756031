"""mxdetection/core (/root/reference/README.md:15-17): anchor and bbox leaf functions of the hot path."""
from . import anchor, bbox, mask  # noqa: F401
