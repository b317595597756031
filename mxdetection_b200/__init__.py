"""mxdetection_b200 - B200-native (sm_100a) implementation of the detection
hot path of jiangzhengkai/mxdetection, behind that toolbox's module layout
(/root/reference/README.md:11-34):

    ops                     RoIAlign forward/backward, NMS          (README.md:24)
    core.anchor             anchor generation, flags, assignment    (README.md:16)
    core.bbox               IoU, delta encode/decode, assigner      (README.md:17)
    models.roi_extractors   FPN level mapping + multi-level RoIAlign (README.md:32)
    models.rpn_heads        proposal stage                          (README.md:28)

Everything executes in hand-written CUDA kernels of libmxdet_sm100.so through
a ctypes C ABI (include/mxdet.h); importing the package without the built
library raises - there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from ._lib import MXDetError, launch_count  # noqa: F401
from . import ops, core, models, parallel  # noqa: F401

__version__ = "0.1.0"
