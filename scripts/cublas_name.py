import torch
from torch.profiler import profile, ProfilerActivity
torch.backends.cuda.matmul.allow_tf32 = True
a = torch.randn(28160, 1152, device="cuda"); b = torch.randn(1152, 1536, device="cuda")
a16, b16 = a.bfloat16(), b.bfloat16()
for _ in range(3): a @ b; a16 @ b16
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    a @ b; a16 @ b16
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(e.key[:150], e.device_time_total)
