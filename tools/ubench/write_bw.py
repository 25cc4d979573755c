import torch, time
dev = torch.device("cuda")
n = 16384 * 50000 * 9
x = torch.empty(n, dtype=torch.float64, device=dev)
for f, name in ((lambda: x.zero_(), "zero_ (fill kernel)"), (lambda: torch.cuda.current_stream().synchronize() or torch.cuda.memset if False else x.fill_(1.0), "fill_(1.0)")):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    print(name, best, "ms", n * 8 / best / 1e6, "GB/s")
# cudaMemsetAsync through cuda-python? use torch's uint8 view zero_
y = x.view(torch.uint8)
import ctypes
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
best = 1e9
for _ in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); cudart.cudaMemsetAsync(x.data_ptr(), 0, n * 8, torch.cuda.current_stream().cuda_stream); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
print("cudaMemsetAsync", best, "ms", n * 8 / best / 1e6, "GB/s")
