"""Microbenchmark: the generator's ResBlock convolutions (B = 16, stage lengths 8192 / 4096 / 2048 / 256 rows at 32 / 64 /
128 / 256 channels, kernels 3 / 7 / 11) with one activation fetch per tap (default) vs one fetch with halo rows
(XVA_GEMM_HALO flag), forward with the leaky-ReLU epilogue and input gradient with the gate. 20 launches per CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (2 * n) * 1e3


print("ch  k  dil rows | fwd per-tap  fwd halo | dgrad per-tap  dgrad halo   (us per launch)")
for ch, T in ((32, 8192), (64, 4096), (128, 2048), (256, 256)):
    x = r(16, T, ch)
    out = torch.empty_like(x)
    for k in (3, 7, 11):
        for dil in (1, 5):
            half = (k - 1) // 2
            shifts = tuple((j - half) * dil for j in range(k))
            w = r(k, ch, ch) * 0.05
            b = r(ch)
            res = []
            for halo in (False, True):
                res.append(timeit(lambda: ops.conv_fwd(x, w, shifts, out=out, bias=b, act_slope=0.1, round_out=True, halo=halo)))
            for halo in (False, True):
                res.append(timeit(lambda: ops.conv_dgrad(x, w, shifts, out=out, gate=x, gate_slope=0.1, round_out=True, halo=halo)))
            print(f"{ch:3d} {k:2d} {dil:3d} {T:5d} | {res[0]:8.1f} {res[1]:9.1f} | {res[2]:9.1f} {res[3]:10.1f}")
