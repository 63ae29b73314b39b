"""Bring-up probe (not a test): tries MN-major descriptor constants for the tap-GEMM and prints the error of a
dgrad (mode 1) and a wgrad (mode 2) case against torch. Usage on the GPU box: python tests/gpu_mn_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__ as ge

ge.build()
from xva_trainer_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


g = torch.Generator(device="cuda").manual_seed(0)
B, T, K, N = 2, 200, 128, 256
dy = torch.randn(B, T, N, device="cuda", generator=g)
w = torch.randn(1, N, K, device="cuda", generator=g) * N ** -0.5
x = torch.randn(B, T, K, device="cuda", generator=g)
want_dx = dy @ w[0]
want_dw = torch.einsum("btn,btk->nk", dy, x)[None]
combos = [None, "1:4096:512:4", "1:4096:1024:4", "1:512:4096:4", "1:4096:512:3", "2:4096:1024:3", "2:4096:1024:4",
          "1:4096:256:4", "1:1024:512:4", "6:4096:512:4", "1:4096:512:6", "1:4096:512:5"]
for c in combos:
    if c is None:
        os.environ.pop("XVA_MN_DEBUG", None)
    else:
        os.environ["XVA_MN_DEBUG"] = c
    try:
        dx = ops.conv_dgrad(dy, w, (0,))
        dw = ops.conv_wgrad(dy, x, (0,), split=1)
        torch.cuda.synchronize()
        print(f"{str(c):>16}  dgrad rel {rel(dx, want_dx):.3e}  wgrad rel {rel(dw, want_dw):.3e}  "
              f"|dx| {dx.abs().max().item():.3g} |dw| {dw.abs().max().item():.3g}", flush=True)
    except Exception as e:
        print(f"{str(c):>16}  ERROR {str(e)[:200]}", flush=True)
