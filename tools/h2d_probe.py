"""H2D bandwidth from default pinned vs write-combined pinned host memory (decides how bench.py's e2e leg allocates)."""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so")
n = 1 << 30
torch.cuda.init()
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
for flags, name in ((0, "cudaHostAllocDefault"), (4, "cudaHostAllocWriteCombined")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    best = 0
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(0)) == 0
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    print(f"{name:28s} {best:6.2f} GB/s")
    rt.cudaFreeHost(p)
