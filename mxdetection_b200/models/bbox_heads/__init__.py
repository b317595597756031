"""mxdetection/models/bbox_heads (/root/reference/README.md:29): detection post-processing only;
the FC layers are out of scope."""
from .post_process import multiclass_nms, get_det_bboxes  # noqa: F401
