"""World-size-2 gloo test of the N>1 host logic: image sharding and the tail gather."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from mxdetection_b200.parallel import gather_detections, shard_range
        lo, hi = shard_range(6, rank, world)
        rng = np.random.default_rng(100)
        allp = torch.from_numpy(rng.standard_normal((6, 5, 5)).astype(np.float32))
        alln = torch.tensor([5, 0, 3, 1, 5, 2], dtype=torch.int32)
        out, cnt = gather_detections(allp[lo:hi], alln[lo:hi], first_image_id=lo)
        q.put((rank, out.numpy(), cnt.numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # surface the failure instead of hanging the parent on q.get
        q.put((rank, repr(e), None))


def test_shard_range_covers_everything():
    from mxdetection_b200.parallel import shard_range
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gather_detections_world2_gloo():
    from mxdetection_b200.parallel import pack_detections
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(100)
    allp = torch.from_numpy(rng.standard_normal((6, 5, 5)).astype(np.float32))
    alln = torch.tensor([5, 0, 3, 1, 5, 2], dtype=torch.int32)
    expect = pack_detections(allp, alln, 0).numpy()      # == concatenation of the single-process result
    for rank, out, cnt in res:
        assert cnt is not None, out
        assert np.array_equal(out, expect) and cnt.tolist() == alln.tolist()
    assert expect[1, 0, 0] == -1 and expect[0, 4, 0] == 0 and expect[5, 2, 0] == -1
