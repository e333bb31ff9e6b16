"""Probe: host->device bandwidth from torch-pinned memory vs cudaHostAlloc(WriteCombined) (e2e is PCIe-bound)."""
import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so")
n = 178 * 1024 * 1024
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9
print("torch pinned  H2D GB/s", round(bw(lambda: dev.copy_(pin, non_blocking=True)), 2))
for flag, name in ((0, "default"), (4, "write-combined"), (1, "portable")):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flag)) == 0
    ctypes.memset(p, 1, n)
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(n), 1, ctypes.c_void_p(s))
    print("cudaHostAlloc", name, "H2D GB/s", round(bw(f), 2))
    rt.cudaFreeHost(p)
# two streams / two halves concurrently
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): dev[: n // 2].copy_(pin[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): dev[n // 2:].copy_(pin[n // 2:], non_blocking=True)
print("torch pinned, 2 streams H2D GB/s", round(bw(two), 2))
d2h = torch.empty(n, dtype=torch.uint8).pin_memory()
print("D2H GB/s", round(bw(lambda: d2h.copy_(dev, non_blocking=True)), 2))
