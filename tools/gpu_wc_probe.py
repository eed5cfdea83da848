"""H2D bandwidth from default pinned vs write-combined pinned host memory."""
import ctypes as C
import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
n = 64 * 3 * 416 * 416 * 4
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, flags in (("default", 0), ("write-combined", 4), ("portable+wc", 5)):
    p = C.c_void_p()
    err = rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    assert err == 0, err
    C.memset(p, 1, n)
    st = torch.cuda.current_stream().cuda_stream
    def cp():
        rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(st))
    for _ in range(3):
        cp()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cp()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:16s}: {n / ms / 1e6:.1f} GB/s")
    rt.cudaFreeHost(p)
