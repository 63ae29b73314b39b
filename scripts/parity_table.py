#!/usr/bin/env python
"""Prints the measured parity of the product path against the CPU oracle (tests/parity_util.py) as a table:
    python scripts/parity_table.py [--quick] > gpurun_out/r02_parity_table.txt      (on the B200 box)
With baseline/_ref present it appends the reference's OWN reduced-precision error (fp16 AMP, cuDNN TF32) against its
strict-fp32 result, measured on the same GPU (baseline/ref_step.py::amp_error_table)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import parity_util as pu  # noqa: E402


def fp_row(tag, r):
    g = r["grad"]
    fwd = " ".join(f"{k}={v:.2e}" for k, v in r["fwd"].items())
    loss = " ".join(f"{k}={v:.1e}" for k, v in r["loss"].items() if v)
    print(f"{tag:44s} | {fwd}")
    print(f"{'':44s} | losses: {loss}")
    print(f"{'':44s} | grads ({g['n']} tensors): global={g['global']:.2e} median={g['median']:.2e} worst={g['worst']:.2e} ({g['worst_key']})")


def main():
    quick = "--quick" in sys.argv
    torch.set_num_threads(os.cpu_count() or 8)
    print("# Relative L2 error of the product path (tcgen05 kind::tf32, operands rounded to nearest, fp32 accumulate) vs the")
    print("# CPU oracle (fp32), one training step on identical weights and inputs, dropout off. GPU:", torch.cuda.get_device_name(0))
    print("\n## FastPitch (forward tensors | losses | parameter gradients)")
    shapes = [("toy 4x40x150", 4, 40, 150)] + ([] if quick else [("BASELINE 32x160x880", 32, 160, 880)])
    for name, B, Tt, Tm in shapes:
        for stage in (2, 3, 4):
            for ragged in (False, True):
                t0 = time.time()
                r = pu.fastpitch_step_errors(stage, B, Tt, Tm, ragged, seed=1234 if B == 32 else 11)
                fp_row(f"{name} stage {stage} {'ragged' if ragged else 'full'} ({time.time() - t0:.0f}s)", r)
    print("\n## FastPitch 5-step trajectory (loss per step: engine, oracle, relative difference; weights after 5 LAMB steps)")
    for stage in (3, 4):
        traj, werr = pu.fastpitch_trajectory(stage, 4, 40, 150, True, steps=5)
        print(f"toy 4x40x150 ragged stage {stage}: " + "  ".join(f"[{t['step']}] {t['loss']:.6f} {t['oracle']:.6f} {t['rel']:.1e}" for t in traj))
        print(f"   weights after 5 steps: global={werr['global']:.2e} worst={werr['worst']:.2e} ({werr['worst_key']})")
    if not quick:
        traj, werr = pu.fastpitch_trajectory(3, 32, 160, 880, False, steps=3, seed=1234)
        print("BASELINE 32x160x880 full stage 3 (3 steps): " + "  ".join(f"[{t['step']}] {t['loss']:.6f} {t['oracle']:.6f} {t['rel']:.1e}" for t in traj))
        print(f"   weights after 3 steps: global={werr['global']:.2e} worst={werr['worst']:.2e} ({werr['worst_key']})")
    print("\n## HiFi-GAN (losses per step | D-step gradients | G-step gradients | weights after the last step)")
    for name, B, frames, steps in [("toy 2x2048", 2, 8, 3)] + ([] if quick else [("BASELINE 16x8192", 16, 32, 1)]):
        t0 = time.time()
        r = pu.hifigan_step_errors(B, frames, steps=steps)
        for s, l in enumerate(r["loss"]):
            print(f"{name} step {s}: " + " ".join(f"{k}={v:.1e}" for k, v in l.items()))
        for kind in ("dgrad", "ggrad"):
            g = r[kind]
            print(f"   {kind} ({g['n']} tensors): global={g['global']:.2e} median={g['median']:.2e} worst={g['worst']:.2e} ({g['worst_key']})")
        for n, w in r["weights"].items():
            print(f"   weights {n}: global={w['global']:.2e} worst={w['worst']:.2e} ({w['worst_key']})")
        print(f"   ({time.time() - t0:.0f}s)")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_step

    if ref_step.available():
        print("\n## The reference's own reduced-precision modes vs its own strict fp32 (unmodified reference modules, PyTorch eager,")
        print("## same GPU; fp32 = cuDNN TF32 convolutions allowed = torch default; amp_fp16 = the trainer's default mode)")
        for shape in ((4, 40, 150),) + (() if quick else ((8, 160, 880),)):
            for stage in (3, 4):
                t = ref_step.amp_error_table("cuda:0", *shape, stage=stage)
                for mode, e in t.items():
                    print(f"reference B{shape[0]}x{shape[1]}x{shape[2]} stage {stage} {mode:9s}: mel_out={e['mel_out']:.2e} loss={e['loss']:.1e} "
                          f"pitch_pred={e['pitch_pred']:.2e} grads: global={e['grad_global']:.2e} median={e['grad_median']:.2e} "
                          f"worst={e['grad_worst'][1]:.2e} ({e['grad_worst'][0]})")
    else:
        print("\n(baseline/_ref absent: the reference's own AMP error was not measured in this run)")


if __name__ == "__main__":
    main()
