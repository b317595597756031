"""mxdetection/ops (/root/reference/README.md:24): RoIAlign fwd/bwd and NMS."""
from .roi_align import (roi_align_forward, roi_align_backward, ROIAlign, RoIAlign, RoIAlignFunction,  # noqa: F401
                        roi_align_fpn_forward, roi_align_fpn_backward)
from .nms import nms, nms_indices, nms_batched, box_nms, box_nms_backward, topk_stable  # noqa: F401
from .host_stage import HostRoIStage  # noqa: F401
