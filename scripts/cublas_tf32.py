import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
def bench(M, N, K, dtype, iters=20):
    a = torch.randn(M, K, device="cuda", dtype=dtype); b = torch.randn(K, N, device="cuda", dtype=dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{str(dtype):16s} {M}x{N}x{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s")
for shape in ((8192, 8192, 8192), (28160, 1536, 1152), (28160, 384, 4608), (28160, 1536, 384)):
    bench(*shape, torch.float32)
    bench(*shape, torch.bfloat16)
