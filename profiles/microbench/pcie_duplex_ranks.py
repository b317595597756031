"""PCIe / host-memory ceiling with ALL ranks copying at once: the denominator of bench.py's e2e number at N GPUs.
Every rank moves the e2e step payload (937 MB H2D + 937 MB D2H, 8 chunks each way, two streams) between its own
pinned buffers and its own GPU, all ranks started by one barrier; the per-rank time is reduced with MAX.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      profiles/microbench/pcie_duplex_ranks.py
prints one JSON line on rank 0: per-rank GB/s each way (slowest rank) and the box aggregate, alone and duplex."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        import bench
        bench.bind_to_gpu_numa_node(local)          # same placement as the bench's e2e leg
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("numa binding skipped: %r\n" % (e,))
    n = 937 * 1000 * 1000 // 4
    chunks = 8
    c = n // chunks
    h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    h_in.zero_(); h_out.zero_()
    d_in = torch.empty(n, dtype=torch.float32, device=dev); d_out = torch.zeros(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d():
        for i in range(chunks):
            d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)

    def d2h():
        for i in range(chunks):
            h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            h2d()
        with torch.cuda.stream(s2):
            d2h()
        cur.wait_stream(s1); cur.wait_stream(s2)

    def run(fn, reps=4):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return min(ts)

    res = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("duplex", both)):
        ms = run(fn)
        res[name] = {"ms_max_over_ranks": round(ms, 3), "per_rank_gbs_each_way": round(n * 4 / ms / 1e6, 2),
                     "box_aggregate_gbs_each_way": round(world * n * 4 / ms / 1e6, 1)}
    if rank == 0:
        print(json.dumps({"ranks": world, "payload_mb_each_way": 937, "chunks": chunks, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
