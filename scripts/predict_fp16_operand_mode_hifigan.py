"""The 16-bit operand what-if of scripts/predict_fp16_operand_mode.py for the HiFi-GAN half: generator forward + backward at a
toy shape (1 x 80 x 6 mel frames -> 1536 samples, all 72 weight-normed convolutions) with GEMM operands stored / read as tf32
(today), fp16 and bf16, against the fp32 oracle + autograd, on the CPU with the emulated C ABI's operand model; and the
magnitude range of every GEMM operand.
    python scripts/predict_fp16_operand_mode_hifigan.py  ->  appended to profiles/r02_fp16_operand_mode_prediction.txt"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cabi_emu  # noqa: E402
from oracle import hifigan as ohg  # noqa: E402
from test_cabi_emu_cpu import HG_PATCHES, rel  # noqa: E402


class H(dict):
    __getattr__ = dict.__getitem__


def main():
    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
    sd = ohg.make_generator_state(3, scale=0.7)
    gen = torch.Generator().manual_seed(3)
    mel = torch.randn(1, 80, 6, generator=gen)
    w = torch.randn(1, 1, 256 * 6, generator=gen)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ohg.generator(leaves, mel)
    (want * w).sum().backward()
    lines = ["", "HiFi-GAN generator, forward + backward at 1 x 80 x 6 frames (scripts/predict_fp16_operand_mode_hifigan.py):",
             f"{'operands':<10} {'grad scale':>10} | {'waveform':>9} | {'grad all':>9} {'median':>9} {'worst':>9}  finite  worst tensor"]
    rng = None
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        for operand, scales in ((None, (1,)), ("fp16", (1, 2 ** 10)), ("bf16", (1,))):
            for sc in scales:
                stats = {} if operand is None else None
                G = hg.Generator(h, device="cpu")
                G.load_state_dict(sd)
                G.train()
                cabi_emu.TF32, cabi_emu.OPERAND16, cabi_emu.OPERAND_STATS = True, operand, stats
                try:
                    y = G(mel)
                    G.zero_grad()
                    G.backward(w * float(sc))
                finally:
                    cabi_emu.TF32, cabi_emu.OPERAND16, cabi_emu.OPERAND_STATS = False, None, None
                grads = {k: p.grad.detach() / float(sc) for k, p in G.named_parameters()}
                per = sorted(((rel(grads[k], leaves[k].grad), k) for k in grads), reverse=True)
                num = sum(float((grads[k] - leaves[k].grad).norm()) ** 2 for k in grads)
                den = sum(float(leaves[k].grad.norm()) ** 2 for k in grads)
                finite = all(bool(torch.isfinite(v).all()) for v in grads.values())
                lines.append(f"{operand or 'tf32':<10} {sc:>10} | {rel(y, want.detach()):9.2e} | {math.sqrt(num / den):9.2e} {per[len(per) // 2][0]:9.2e} "
                             f"{per[0][0]:9.2e}  {str(finite):<6}  {per[0][1]}")
                print(lines[-1], flush=True)
                if stats is not None:
                    rng = stats
    lines += ["", f"{'launch kind':<34} {'elements':>12} {'min |x|':>10} {'max |x|':>10} {'< 6.1e-5':>9} {'< 6.0e-8':>9} {'> 65504':>8}"]
    names = {0: "forward (mode 0)", 1: "input gradient (mode 1)", 2: "weight gradient (mode 2)"}
    for (mode, which), st in sorted(rng.items()):
        lines.append(f"{names[mode] + ', operand ' + which:<34} {st['n']:>12} {st['min']:10.1e} {st['max']:10.1e} "
                     f"{st['below_normal'] / st['n']:9.1%} {st['below_subnormal'] / st['n']:9.1%} {st['over']:>8}")
    with open(os.path.join(ROOT, "profiles", "r02_fp16_operand_mode_prediction.txt"), "a") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    main()
