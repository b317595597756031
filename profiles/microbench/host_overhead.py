import sys, os, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import torch
from mxdetection_b200 import synthetic as syn
from mxdetection_b200.models.rpn_heads import RPNHead, ProposalConfig
B = 2
dev = "cuda"
r = syn.rpn_inputs(2, B, 800, 1088)
sc = [torch.from_numpy(s).to(dev) for s in r["scores"]]; dl = [torch.from_numpy(x).to(dev) for x in r["deltas"]]
shp = torch.from_numpy(r["img_shapes"]).to(dev)
head, cfg = RPNHead(), ProposalConfig(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7)
fn = lambda: head.get_proposals(sc, dl, r["feat_shapes"], shp, cfg)
for _ in range(5): fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): fn()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue us/call", (t1 - t0) / 200 * 1e6, "total us/call", (t2 - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fn()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
