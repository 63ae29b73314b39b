"""Where do the ~23 us of a small tap-GEMM launch go? In-kernel cycle counters of CTA 0 (XVA_GEMM_DBG=32: MMA warp waiting
for an accumulator / for operands / issuing; epilogue warp waiting / working; total) next to the CUDA-event time per launch,
plus the time of an empty kernel launched the same way (launch-to-launch floor of the stream)."""
import ctypes as C, os, sys
os.environ["XVA_GEMM_DBG"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import capi, ops


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


def timeit(run, n=30):
    for _ in range(5):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


clk = 1.965e3  # MHz -> cycles per us (max clock; the chip is not power-limited on these launches)
for B, T in ((32, 160), (32, 880)):
    x, vec = gen(B, T, 384, seed=1), gen(B, T, 64, seed=2)
    wq, bq, wo = gen(1, 192, 384, seed=4, scale=0.05), gen(192, seed=5), gen(1, 384, 64, seed=6, scale=0.1)
    out_q = torch.empty(B, T, 192, device="cuda")
    out_o = torch.empty(B, T, 384, device="cuda")
    for name, run in (("qkv fwd K=384 N=192", lambda: ops.conv_fwd(x, wq, bias=bq, round_out=True, out=out_q)),
                      ("o_net fwd K=64 N=384 +res", lambda: ops.conv_fwd(vec, wo, residual=x, out=out_o))):
        us = timeit(run)
        cnt = (C.c_longlong * 8)()
        capi.load().xva_gemm_debug_counters(C.byref(cnt))
        c = list(cnt)
        print(f"B={B} T={T} {name:28s} {us:6.1f} us/launch | CTA0: tiles {c[5]} total {c[6] / clk:5.1f} us = MMA warp: wait-acc {c[0] / clk:4.1f} "
              f"wait-operands {c[1] / clk:4.1f} issue {c[2] / clk:4.1f} | epilogue warp: wait {c[3] / clk:4.1f} work {c[4] / clk:4.1f}", flush=True)
z = torch.zeros(1, device="cuda")
print(f"torch fill_ of 1 float (an almost empty kernel), back to back: {timeit(lambda: z.fill_(1.0)):.1f} us/launch")
y = torch.empty_like(x)
print(f"torch copy of 43 MB (32x880x384 fp32) , back to back:       {timeit(lambda: y.copy_(x)):.1f} us/launch")
