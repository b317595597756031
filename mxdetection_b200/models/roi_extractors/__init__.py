"""mxdetection/models/roi_extractors (/root/reference/README.md:32)."""
from .single_level import SingleLevelRoI, map_roi_levels  # noqa: F401
