"""How fast is a strided (column-block) host-to-device copy? cudaMemcpy2DAsync of `cols` of the 14 columns of a pinned
row-major 2^20 x 14 u64 matrix, against the contiguous copy of the whole matrix."""
import ctypes as C
import sys
import torch

rt = C.CDLL("libcudart.so.12") if len(sys.argv) < 2 else C.CDLL(sys.argv[1])
n, w = 1 << 20, 14
host = torch.empty((n, w), dtype=torch.int64).pin_memory()
host.random_()
dev = torch.empty((n, w), dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    return best


t = timed(lambda: rt.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), n * w * 8, 1, st))
print("contiguous %d MB: %.3f ms = %.1f GB/s" % (n * w * 8 >> 20, t, n * w * 8 / t / 1e6))
for cols in (7, 4, 2, 1):
    # column block -> its own dense n x cols device matrix
    t = timed(lambda: rt.cudaMemcpy2DAsync(dev.data_ptr(), cols * 8, host.data_ptr(), w * 8, cols * 8, n, 1, st))
    print("2D %2d of 14 columns (%3d B per row): %.3f ms = %.1f GB/s" % (cols, cols * 8, t, n * cols * 8 / t / 1e6))
