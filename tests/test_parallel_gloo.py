"""World-size-2 gloo test of the N>1 host logic: image sharding and the tail gather."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def pack_reference(proposals, num_valid, first_image_id):
    """Layout of mxd_pack_detections restated on the CPU: (b,m,5) + (b) -> (b,m+1,6)."""
    b, m, _ = proposals.shape
    out = torch.zeros((b, m + 1, 6), dtype=torch.float32)
    for i in range(b):
        nv = int(num_valid[i])
        out[i, 0, 0] = nv; out[i, 0, 1] = first_image_id + i
        out[i, 1:, 1:] = proposals[i]
        out[i, 1:, 0] = -1.0
        out[i, 1:1 + nv, 0] = first_image_id + i
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from mxdetection_b200.parallel import all_gather_packed, shard_range, unpack_detections
        lo, hi = shard_range(6, rank, world)
        rng = np.random.default_rng(100)
        allp = torch.from_numpy(rng.standard_normal((6, 5, 5)).astype(np.float32))
        alln = torch.tensor([5, 0, 3, 1, 5, 2], dtype=torch.int32)
        # the packing kernel needs a GPU (tests/test_gpu_parity.py checks it against this layout); here the host logic:
        # shard -> one all-gather of the packed blocks -> views
        out, cnt, ids = unpack_detections(all_gather_packed(pack_reference(allp[lo:hi], alln[lo:hi], lo)))
        q.put((rank, out.numpy(), cnt.numpy().astype(np.int32)))
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
    expect = pack_reference(allp, alln, 0).numpy()[:, 1:]      # == concatenation of the single-process result
    for rank, out, cnt in res:
        assert cnt is not None, out
        assert np.array_equal(out, expect) and cnt.tolist() == alln.tolist()
    assert expect[1, 0, 0] == -1 and expect[0, 4, 0] == 0 and expect[5, 2, 0] == -1
