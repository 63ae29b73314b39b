"""Diagnostic: per-parameter error of the xVAPitch decoder's backward with the exact-fp32 checker GEMM (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import hifigan as ohg  # noqa: E402
from xva_trainer_b200 import capi, ops  # noqa: E402
import test_hifigan_gpu as T  # noqa: E402

capi.load()
orig = ops.gemm_launch
ops.gemm_launch = lambda args, ref=False: orig(args, True)
capi.call("xva_set_operand_rounding", 0)
gold, spec, sd, z, cond = T._xvapitch_fixture()
for with_cond in (True, False):
    for dbl in (False, True):
        d = T._xvapitch_decoder(None, sd)
        w = torch.randn(2, 1, 1536, generator=torch.Generator().manual_seed(22))
        y = d(z.cuda(), g=cond.cuda() if with_cond else None)
        d.zero_grad()
        dz, dg = d.backward(w.cuda(), need_input_grad=True)
        cast = (lambda t: t.double()) if dbl else (lambda t: t.clone())
        leaves = {k: cast(v).requires_grad_(True) for k, v in sd.items()}
        zl, gl = cast(z).requires_grad_(True), cast(cond).requires_grad_(True)
        yo = ohg.generator_vits(leaves, zl, gl if with_cond else None)
        (yo * cast(w)).sum().backward()
        errs = sorted(((T.rel(p.grad, leaves[k].grad), k) for k, p in d.named_parameters() if leaves[k].grad is not None
                       and p.grad is not None), reverse=True)
        print(f"cond={with_cond} oracle={'f64' if dbl else 'f32'} fwd {T.rel(y, yo):.2e} dz {T.rel(dz, zl.grad):.2e}",
              "dg %.2e" % T.rel(dg, gl.grad) if with_cond else "", "worst:", [(f"{e:.1e}", k) for e, k in errs[:6]],
              "median %.1e" % errs[len(errs) // 2][0])
