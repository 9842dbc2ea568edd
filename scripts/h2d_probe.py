"""H2D bandwidth probe: default pinned vs write-combined pinned memory, one vs two copy streams (205 MB, as one bench step)."""
import ctypes, time
import numpy as np, torch
rt = ctypes.CDLL("libcudart.so.12")
N = 204_800_000
dev = torch.device("cuda", 0)
dst = torch.empty(N, dtype=torch.uint8, device=dev)
def alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(flags))
    assert rc == 0, rc
    ctypes.memset(p, 1, N)
    return p
def bench(p, nstreams, reps=10):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    part = N // nstreams
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, s in enumerate(streams):
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + k * part), ctypes.c_void_p(p.value + k * part), ctypes.c_size_t(part), 1, ctypes.c_void_p(s.cuda_stream))
            assert rc == 0
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return N / best / 1e9
for name, flags in (("default pinned", 0), ("portable", 1), ("write-combined", 4)):
    p = alloc(flags)
    print("%-16s 1 stream %.1f GB/s | 2 streams %.1f GB/s | 4 streams %.1f GB/s" % (name, bench(p, 1), bench(p, 2), bench(p, 4)), flush=True)
    rt.cudaFreeHost(p)
t = torch.empty(N, dtype=torch.uint8).pin_memory()
p = ctypes.c_void_p(t.data_ptr())
print("%-16s 1 stream %.1f GB/s | 2 streams %.1f GB/s" % ("torch pin_memory", bench(p, 1), bench(p, 2)))
# D2H for completeness
src = torch.empty(N // 8, dtype=torch.uint8, device=dev)
