"""One ReLU gate is enough: the FastPitch step at the toy shape of the parity table (4 x 40 x 150 ragged, stage 3) run twice on
the CPU through the emulated C ABI in EXACT arithmetic -- once with the fused attention entry points, once with the six-launch
chain. The two forwards agree to 5e-7, yet one pre-activation of the energy predictor's second ConvReLUNorm layer is -0.0 in
one run and +1.9e-7 in the other, its ReLU gate decides differently, and that single gate (of 40 960 in the layer) moves the
gradient vector by 1.6e-3 and pitch_emb.weight by 7.5e-3 against the oracle. Prints the flipped gates per predictor layer.
    python scripts/diag_relu_gate_flip.py"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cabi_emu
from oracle import fastpitch as ofp
from test_cabi_emu_cpu import FP_PATCHES
x, y = ofp.synthetic_batch(4, 40, 150, seed=11, ragged=True)
sd = ofp.make_state(1234)
pre = {}
with cabi_emu.installed():
    fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
    for fused in (True, False):
        m = fp.FastPitch(device="cpu"); m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.training_stage = 3; m.train(); m.p_drop = 0.0; m.fused_attn = fused
        o = m(x)
        pre[fused] = {n: (m._ctx.preds[n]["s1"]["pre"].clone(), m._ctx.preds[n]["s2"]["pre"].clone()) for n in ("pitch", "energy")}
for n in ("pitch", "energy"):
    for li in (0, 1):
        a, b = pre[True][n][li], pre[False][n][li]
        flips = ((a > 0) != (b > 0))
        print(n, li, "flips:", int(flips.sum()), "values at flips:", a[flips].tolist()[:4], b[flips].tolist()[:4])
