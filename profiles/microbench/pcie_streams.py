import torch
dev = torch.device("cuda:0")
n = 937 * 1000 * 1000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev); d_out = torch.empty(n, dtype=torch.float32, device=dev)
def run(chunks, nst):
    si = [torch.cuda.Stream(dev) for _ in range(nst)]; so = [torch.cuda.Stream(dev) for _ in range(nst)]
    c = n // chunks
    def both():
        cur = torch.cuda.current_stream()
        for s in si + so: s.wait_stream(cur)
        for i in range(chunks):
            with torch.cuda.stream(si[i % nst]): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(so[i % nst]): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
        for s in si + so: cur.wait_stream(s)
    both(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); both(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("chunks %d, %d stream(s) per direction: %.2f ms (%.1f GB/s each way)" % (chunks, nst, best, n * 4 / best / 1e6))
for chunks, nst in ((1, 1), (8, 1), (8, 2), (8, 4), (32, 2), (32, 4), (2, 2)):
    run(chunks, nst)
