"""Does a WRITE-COMBINED pinned source buffer speed up the H2D half of the e2e step (review item, round 1)?
H2D alone and duplex (H2D from normal / write-combined pinned memory || D2H into normal pinned memory), 937 MB each way
in 8 chunks, the e2e step's payload.   python profiles/microbench/pcie_wc.py"""
import ctypes
import json

import torch

rt = ctypes.CDLL("libcudart.so.12")
N = 937 * 1000 * 1000
CH = 8


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p


def main():
    torch.cuda.init()
    dev = torch.device("cuda:0")
    d_in = torch.empty(N, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(N, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs = {"pinned": host_alloc(N, 0), "write_combined": host_alloc(N, 4)}       # cudaHostAllocWriteCombined = 4
    h_out = host_alloc(N, 0)
    for p in list(bufs.values()) + [h_out]:
        ctypes.memset(p, 1, N)
    c = N // CH

    def h2d(src, st):
        for i in range(CH):
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr() + i * c), ctypes.c_void_p(src.value + i * c), ctypes.c_size_t(c),
                                    ctypes.c_int(1), ctypes.c_void_p(st.cuda_stream))
            assert rc == 0

    def d2h(st):
        for i in range(CH):
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(h_out.value + i * c), ctypes.c_void_p(d_out.data_ptr() + i * c), ctypes.c_size_t(c),
                                    ctypes.c_int(2), ctypes.c_void_p(st.cuda_stream))
            assert rc == 0

    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    res = {}
    cur = torch.cuda.current_stream()
    for name, src in bufs.items():
        def alone():
            s1.wait_stream(cur); h2d(src, s1); cur.wait_stream(s1)

        def both():
            s1.wait_stream(cur); s2.wait_stream(cur)
            h2d(src, s1); d2h(s2)
            cur.wait_stream(s1); cur.wait_stream(s2)
        a, b = timed(alone), timed(both)
        res[name] = {"h2d_ms": round(a, 2), "h2d_gbs": round(N / a / 1e6, 1), "duplex_ms": round(b, 2), "duplex_gbs_each_way": round(N / b / 1e6, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
