"""How fast is a cudaMemsetAsync (copy engine / driver fill) of a full-resolution gradient map, eagerly and as a node of a
CUDA graph, next to a fill kernel (torch zero_)?  usage: python tools/ubench/memset_rate.py"""
import ctypes
import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
dev = torch.device("cuda:0")
n = 8 * 21 * 448 * 448
x = torch.empty(n, dtype=torch.float32, device=dev)
nbytes = n * 4
s = torch.cuda.Stream()


def timed(fn, reps=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def memset():
    rt.cudaMemsetAsync(x.data_ptr(), 0, nbytes, s.cuda_stream)


print(f"{nbytes / 1e6:.0f} MB")
t = timed(memset)
print(f"cudaMemsetAsync eager : {t:7.1f} us  {nbytes / t / 1e6:.2f} TB/s")
t = timed(lambda: x.zero_())
print(f"torch zero_ (kernel)  : {t:7.1f} us  {nbytes / t / 1e6:.2f} TB/s")
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(10):
            memset()
t = timed(lambda: g.replay(), reps=5) / 10
print(f"memset node in a graph: {t:7.1f} us  {nbytes / t / 1e6:.2f} TB/s")
