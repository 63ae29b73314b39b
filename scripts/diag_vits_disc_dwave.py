"""Diagnostic: dL/d(waveform) through each net of the VITS discriminator vs autograd through the oracle, on the tf32 path
and with the exact-fp32 checker GEMM (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_hifigan_gpu as T
from oracle import hifigan as ohg
from xva_trainer_b200 import capi, ops

hg, m, gold, spec, sd, x, x_hat = T._vits_disc(None)
B, Tn = x.shape[0], x.shape[2]
# oracle per-net gradients
xh = x_hat.clone().requires_grad_(True)
sr, fr, sg, fg = ohg.vits_discriminator(sd, x, xh)
want_gen, want_fm = [], []
for i in range(6):
    (g1,) = torch.autograd.grad(ohg.generator_loss([sg[i]]), xh, retain_graph=True)
    (g2,) = torch.autograd.grad(ohg.feature_loss([fr[i]], [fg[i]]), xh, retain_graph=True)
    want_gen.append(g1.reshape(B, Tn)); want_fm.append(g2.reshape(B, Tn))
orig = ops.gemm_launch
for exact in (False, True):
    if exact:
        ops.gemm_launch = lambda args, ref=False: orig(args, True)
        capi.call("xva_set_operand_rounding", 0)
    for fm in (False, True):
        errs = []
        for i in range(6):
            xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
            # keep only net i: zero the others' contribution by running the loss on a one-net view of the model
            class One:
                pass
            one = One(); one.discriminators = [m.discriminators[i]]; one._ctx = [m._ctx[i]]; one._packer = m._packer; one._branch_inputs = None
            dwave = torch.zeros(B, Tn, device="cuda")
            hg.generator_adv_loss_backward(one, [hs[i]], [xf[i]], [hf[i]], dwave, pools=0, fm_grad=fm)
            want = want_gen[i] + (want_fm[i] if fm else 0)
            errs.append(T.rel(dwave.cpu(), want))
        print("exact" if exact else "tf32 ", "fm_grad", fm, "per-net rel err:", " ".join(f"{e:.4f}" for e in errs))
