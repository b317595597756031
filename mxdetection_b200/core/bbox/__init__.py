"""mxdetection/core/bbox (/root/reference/README.md:17)."""
from .geometry import bbox_overlaps  # noqa: F401
from .transforms import bbox2delta, delta2bbox, bbox2roi  # noqa: F401
from .assignment import MaxIoUAssigner, AssignResult, bbox_assign  # noqa: F401
from .sampling import RandomSampler, SamplingResult, pack_targets  # noqa: F401
