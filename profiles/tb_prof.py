"""Cycles the tile backward spends blocked / working (library built with NVCC_EXTRA=-DMXD_TB_PROF):
   python profiles/tb_prof.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mxdetection_b200 import _lib as L  # noqa: E402
from mxdetection_b200 import synthetic as syn  # noqa: E402
from mxdetection_b200.ops import roi_align_fpn_backward  # noqa: E402

dev = "cuda"
d = syn.cfg3(batch=8, with_features=False)
shapes = [(8, 256, h, w) for h, w in d["feat_shapes"]]
rois = torch.from_numpy(d["rois"]).to(dev)
gout = torch.randn((rois.shape[0], 256, 7, 7), device=dev)
grads = [torch.empty(s, device=dev) for s in shapes]
for _ in range(3):
    roi_align_fpn_backward(gout, rois, shapes, (7, 7), d["scales"], 2, grad_feats=grads)
torch.cuda.synchronize()
ws = [v for k, v in L._WORKSPACES.items() if k[2] == "roi_align"][0]
buf = (ctypes.c_ulonglong * 6)()
L.lib.mxd_tb_prof(ctypes.c_void_p(ws.data_ptr()), buf)
p_emp, p_all, c_full, c_row, c_wr, c_all = [int(x) for x in buf]
print("producer: blocked on a free stage %.1f%% of %.0f kcyc | consumer warps: waiting for messages %.1f%%, accumulating rows %.1f%%, "
      "writing tiles %.1f%% of %.0f kcyc" % (100.0 * p_emp / p_all, p_all / 148 / 1e3, 100.0 * c_full / c_all, 100.0 * c_row / c_all,
                                             100.0 * c_wr / c_all, c_all / 148 / 31 / 1e3))
