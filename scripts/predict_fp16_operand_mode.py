"""What would a 16-bit operand mode (tcgen05 kind::f16, twice the tf32 MMA rate, half the operand bytes -- DESIGN.md section 7
#4) do to parity? Answered on the CPU with the emulated C ABI's operand model (tests/cabi_emu.py): the
FastPitch training step at the toy shape of the parity table (4 x 40 x 150 ragged, stage 3) with GEMM operands stored / read
as tf32 (today), fp16 and bf16, with and without a loss scale on the backward pass, against the fp32 oracle; plus the magnitude
range of every GEMM operand (forward = mode 0, input gradient = mode 1, weight gradient = mode 2).
    python scripts/predict_fp16_operand_mode.py  ->  profiles/r02_fp16_operand_mode_prediction.txt"""
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


def run(fp, operand, loss_scale, stats=None, stage=3):
    x, y = ofp.synthetic_batch(4, 40, 150, seed=11, ragged=True)
    sd = ofp.make_state(1234)
    m = fp.FastPitch(device="cpu")
    m.training_stage = stage
    m.train()
    m.p_drop = 0.0
    m.fused_attn = False          # the attention products as GEMM launches too, so their operands are modelled and counted
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    cabi_emu.TF32, cabi_emu.OPERAND16, cabi_emu.OPERAND_STATS = True, operand, stats
    try:
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        o = m(x)
        loss, meta = crit(o, y)
        m.zero_grad()
        m.backward(crit, float(loss_scale))
    finally:
        cabi_emu.TF32, cabi_emu.OPERAND16, cabi_emu.OPERAND_STATS = False, None, None
    want = ofp.forward(sd, x, stage)
    wmeta, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, stage, 1e-3, {}, drop=0.0, training=False)
    got = {k: v / float(loss_scale) for k, v in m.grads(fp.trainable_keys(stage)).items()}
    gs = grad_summary(got, wgrads)
    fwd = {n: rel(g_.float(), w_.float()) for n, g_, w_ in zip(FWD_NAMES, o[:8], want[:8])
           if w_ is not None and w_.dtype != torch.bool and n in ("mel_out", "pitch_pred", "energy_pred")}
    finite = all(bool(torch.isfinite(v).all()) for v in got.values())
    return fwd, gs, finite


def main():
    lines = [__doc__.split("\n    python")[0], ""]
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        lines.append(f"{'operands':<10} {'loss scale':>10} | {'mel_out':>9} {'pitch':>9} {'energy':>9} | {'grad all':>9} {'median':>9} {'worst':>9}  finite  worst tensor")
        for operand, scales in ((None, (1,)), ("fp16", (1, 2 ** 8, 2 ** 12, 2 ** 16)), ("bf16", (1,))):
            for sc in scales:
                stats = {} if (operand is None and sc == 1) else None
                fwd, gs, finite = run(fp, operand, sc, stats)
                lines.append(f"{operand or 'tf32':<10} {sc:>10} | {fwd['mel_out']:9.2e} {fwd['pitch_pred']:9.2e} {fwd['energy_pred']:9.2e} | "
                             f"{gs['global']:9.2e} {gs['median']:9.2e} {gs['worst']:9.2e}  {str(finite):<6}  {gs['worst_key']}")
                print(lines[-1], flush=True)
                if stats is not None:
                    rng = stats
        lines += ["", "Magnitude range of the non-zero GEMM operand elements of that step (fp32 values as today's kernels hand them over;",
                  "fp16: smallest normal 6.1e-5, smallest subnormal 6.0e-8, largest 65504):",
                  f"{'launch kind':<34} {'elements':>12} {'min |x|':>10} {'max |x|':>10} {'< 6.1e-5':>9} {'< 6.0e-8':>9} {'> 65504':>8}"]
        names = {0: "forward (mode 0)", 1: "input gradient (mode 1)", 2: "weight gradient (mode 2)"}
        for (mode, which), st in sorted(rng.items()):
            lines.append(f"{names[mode] + ', operand ' + which:<34} {st['n']:>12} {st['min']:10.1e} {st['max']:10.1e} "
                         f"{st['below_normal'] / st['n']:9.1%} {st['below_subnormal'] / st['n']:9.1%} {st['over']:>8}")
    open(os.path.join(ROOT, "profiles", "r02_fp16_operand_mode_prediction.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    main()
