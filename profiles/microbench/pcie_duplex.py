"""Raw PCIe rates of the box, as the denominator of bench.py's e2e number: pinned H2D alone, D2H alone, and both at
once (two streams), whole 937 MB step payload and per-image 117 MB chunks.  Usage: python profiles/microbench/pcie_duplex.py"""
import torch

def rate(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6, best

def main():
    dev = torch.device("cuda:0")
    for mb, chunks in ((937, 1), (937, 8), (937, 32)):
        n = mb * 1000 * 1000 // 4
        h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
        d_in = torch.empty(n, dtype=torch.float32, device=dev); d_out = torch.empty(n, dtype=torch.float32, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        c = n // chunks
        def h2d():
            for i in range(chunks): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        def d2h():
            for i in range(chunks): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
        def both():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur); s2.wait_stream(cur)
            with torch.cuda.stream(s1): h2d()
            with torch.cuda.stream(s2): d2h()
            cur.wait_stream(s1); cur.wait_stream(s2)
        a, ta = rate(h2d, n * 4); b, tb = rate(d2h, n * 4); c2, tc = rate(both, n * 4)
        print("payload %d MB in %d chunk(s): H2D %.1f GB/s (%.2f ms)  D2H %.1f GB/s (%.2f ms)  duplex %.1f GB/s each way (%.2f ms)"
              % (mb, chunks, a, ta, b, tb, c2, tc))

if __name__ == "__main__":
    main()
