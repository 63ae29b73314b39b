"""Where does the product path's error come from? An ablation on the CPU with the emulated C ABI's tf32 operand model
(tests/cabi_emu.py): the FastPitch training step at the toy shape of the parity table (4 x 40 x 150 ragged, stage 3) with only
ONE class of GEMM operands carried in tf32 at a time -- the weights' operand copy, the operands the forward pass produces
(activations), the operands the backward pass produces (gradients) -- and everything else exact.
    python scripts/tf32_error_budget.py  ->  profiles/r02_tf32_error_budget.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cabi_emu  # noqa: E402
from oracle import fastpitch as ofp  # noqa: E402
from parity_util import FWD_NAMES, grad_summary  # noqa: E402
from test_cabi_emu_cpu import FP_PATCHES, rel  # noqa: E402


def run(fp, what, stage=3):
    x, y = ofp.synthetic_batch(4, 40, 150, seed=11, ragged=True)
    sd = ofp.make_state(1234)
    m = fp.FastPitch(device="cpu")
    m.training_stage = stage
    m.train()
    m.p_drop = 0.0
    m.fused_attn = False
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    cabi_emu.TF32, cabi_emu.ROUND_WHAT, cabi_emu.MMA_TRUNCATES = True, set(what), False
    try:
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        cabi_emu.PHASE = "fwd"
        o = m(x)
        loss, meta = crit(o, y)
        m.zero_grad()
        cabi_emu.PHASE = "bwd"
        m.backward(crit, 1.0)
    finally:
        cabi_emu.TF32, cabi_emu.ROUND_WHAT, cabi_emu.MMA_TRUNCATES, cabi_emu.PHASE = False, {"weights", "fwd", "bwd"}, True, "fwd"
    want = ofp.forward(sd, x, stage)
    wmeta, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, stage, 1e-3, {}, drop=0.0, training=False)
    gs = grad_summary(m.grads(fp.trainable_keys(stage)), wgrads)
    return rel(o[0].float(), want[0].float()), gs


def main():
    lines = [__doc__.split("\n    python")[0], "",
             f"{'operands carried in tf32':<44} {'mel_out':>9} | {'grad all':>9} {'median':>9} {'worst':>9}  worst tensor"]
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        for label, what in (("none (exact)", ()), ("weights only", ("weights",)), ("forward operands (activations) only", ("fwd",)),
                            ("backward operands (gradients) only", ("bwd",)), ("weights + activations", ("weights", "fwd")),
                            ("all three (the product path)", ("weights", "fwd", "bwd"))):
            e, gs = run(fp, what)
            lines.append(f"{label:<44} {e:9.2e} | {gs['global']:9.2e} {gs['median']:9.2e} {gs['worst']:9.2e}  {gs['worst_key']}")
            print(lines[-1], flush=True)
    lines += ["", "(independent rounding errors add in quadrature: sqrt of the sum of the squared single-class rows ~ the last row)"]
    open(os.path.join(ROOT, "profiles", "r02_tf32_error_budget.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
