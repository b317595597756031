"""Small driver for profiling the non-RoIAlign stages: RPN proposals (config 2) and the assigner (config 4b).
python profiles/stage_bench.py [batch] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.models.rpn_heads import RPNHead, ProposalConfig
from mxdetection_b200.core.anchor import AnchorGenerator, anchor_assign, anchor_inside_flags

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
r = syn.rpn_inputs(2, B, 800, 1088)
sc = [torch.from_numpy(s).to(dev) for s in r["scores"]]; dl = [torch.from_numpy(x).to(dev) for x in r["deltas"]]
shp = torch.from_numpy(r["img_shapes"]).to(dev)
head, cfg = RPNHead(), ProposalConfig(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
a = syn.assigner_inputs(4, B)
anchors, valid = [], []
for (fh, fw), s in zip(a["feat_shapes"], a["strides"]):
    ag = AnchorGenerator(s, [8], [0.5, 1.0, 2.0])
    anchors.append(ag.grid_anchors((fh, fw), s)); valid.append(ag.valid_flags((fh, fw), (fh, fw)))
anchors = torch.cat(anchors); inside = anchor_inside_flags(anchors, torch.cat(valid), a["img_shape"], 0)
gts = torch.from_numpy(a["gts"]).to(dev); ngt = torch.from_numpy(a["num_gts"]).to(dev); gl = torch.from_numpy(a["gt_labels"]).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for name, fn in (("proposals", lambda: head.get_proposals(sc, dl, r["feat_shapes"], shp, cfg)),
                 ("assigner", lambda: anchor_assign(anchors, inside, gts, ngt, gl))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = ev(), ev(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(name, "batch", B, "ms", sum(ts) / len(ts))
