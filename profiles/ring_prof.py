"""Cycles the row-ring forward spends blocked on its barriers (library built with -DMXD_RING_PROF):
   NVCC_EXTRA=-DMXD_RING_PROF python mxdetection_b200/build.py --force && python profiles/ring_prof.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mxdetection_b200 import _lib as L  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.models.roi_extractors import map_roi_levels  # noqa: E402
from mxdetection_b200.ops import roi_align_fpn_forward  # noqa: E402

dev = "cuda"
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
feats = [torch.randn(s, device=dev) for s in shapes]
rois_all = torch.from_numpy(d["rois"]).to(dev)
lv = map_roi_levels(rois_all, 4).cpu().numpy()
for name, sel in [("all", np.ones(len(lv), bool))] + [("L%d" % l, lv == l) for l in range(4)]:
    rois = rois_all[torch.from_numpy(np.nonzero(sel)[0]).to(dev)].contiguous()
    out = torch.empty((rois.shape[0], 256, 7, 7), device=dev)
    for _ in range(3):
        roi_align_fpn_forward(feats, rois, (7, 7), d["scales"], 2, out=out)
    torch.cuda.synchronize()
    ws = [v for k, v in L._WORKSPACES.items() if k[2] == "roi_align"][0]
    buf = (ctypes.c_ulonglong * 6)()
    L.lib.mxd_ring_prof(ctypes.c_void_p(ws.data_ptr()), buf)
    p_desc, p_slot, p_all, c_desc, c_data, c_all = [int(x) for x in buf]
    print("%s: producer blocked on descriptors %.1f%%, on slots %.1f%% | consumer warps blocked on descriptors %.1f%%, on data %.1f%% "
          "(producer %.0f kcyc per CTA, consumer %.0f kcyc per warp)" %
          (name, 100.0 * p_desc / p_all, 100.0 * p_slot / p_all, 100.0 * c_desc / c_all, 100.0 * c_data / c_all,
           p_all / 148 / 1e3, c_all / 148 / 1e3))
