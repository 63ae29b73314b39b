"""One full training step at the BASELINE shapes -- FastPitch stage 3 and 4 at 32 x 880 frames x 160 tokens
(full-length and ragged), HiFi-GAN G + MPD + MSD at 16 x 8192 samples -- PRODUCT path (tcgen05 tf32 tap-GEMMs through the C
ABI) against the CPU oracle: every forward tensor, every loss term, the gradient of every parameter. These are the shapes
the bench times; they exercise the tile tails, segmented tiles, CTA pairs and multi-wave schedules the toy shapes of
test_fastpitch_gpu.py / test_hifigan_gpu.py never reach. Plus a 5-step loss trajectory against the oracle.

Bounds are at most 2x what profiles/r02_parity_table.txt records for the same measurement (scripts/parity_table.py prints
it from the same functions, tests/parity_util.py). The north star's 1e-3 holds for mel_out, pitch_pred and every loss term
(losses: <= 5e-5); energy_pred is at 1.3e-3 and the gradient VECTOR at 1.4e-3 (stage 3) / 6e-4 (stage 4), single small
tensors up to 1e-2 -- the same place the unmodified reference's own default arithmetic on this GPU (cuDNN TF32) lands
against strict fp32 (1.6e-3 / 1.3e-2), and 3-10x closer than its fp16-autocast default (numbers below).
"""
import pytest
import torch

import parity_util as pu

pytestmark = pytest.mark.gpu

# Bounds = 2x the value measured on B200 for the same call (profiles/r02_parity_table.txt, rounded up), never looser than
# that; mel_out and every loss are inside the north star's 1e-3. Measured at 32 x 160 x 880 (full / ragged):
#   stage 3: mel_out 8.8e-4 / 8.9e-4, pitch_pred 9.8e-4 / 9.3e-4, energy_pred 1.30e-3 / 1.26e-3 (the one forward tensor
#            above 1e-3: a near-zero-mean predictor output, absolute error 1e-3 of its rms), losses <= 5.2e-5,
#            gradients global 1.35e-3 / 1.46e-3, median 7.4e-4 / 7.8e-4, worst tensor 8.8e-3 / 9.7e-3 (encoder.word_emb)
#   stage 4: losses 8.3e-6 / 9.2e-6, gradients global 5.9e-4 / 6.0e-4, median 6.5e-4 / 6.7e-4, worst 2.5e-3 / 2.8e-3
# For scale, the UNMODIFIED reference under PyTorch eager on the same B200 against its own strict-fp32 run (8 x 160 x 880):
#   torch-default fp32 (cuDNN TF32): mel_out 2.6e-4, gradients global 1.6e-3, worst 1.3e-2
#   fp16 autocast (the trainer's default): mel_out 1.3e-2, pitch_pred 5.3e-3, gradients global 4.7e-3, worst 3.6e-2
FP_BOUNDS = {
    3: dict(mel=1e-3, fwd=2.6e-3, loss=1.1e-4, g_global=3e-3, g_median=1.6e-3, g_worst=2e-2),
    4: dict(mel=1e-3, fwd=2.6e-3, loss=2e-5, g_global=1.2e-3, g_median=1.4e-3, g_worst=6e-3),
}


@pytest.mark.parametrize("stage", [3, 4])
@pytest.mark.parametrize("ragged", [False, True])
def test_fastpitch_step_at_baseline_shape(lib, stage, ragged):
    r = pu.fastpitch_step_errors(stage, 32, 160, 880, ragged, seed=1234)
    b = FP_BOUNDS[stage]
    assert all(r["exact"].values()), r["exact"]                      # dec_mask (integer length-regulator path)
    for n, e in r["fwd"].items():
        tol = {"mel_out": b["mel"], "pitch_tgt": 1e-5, "energy_tgt": 1e-5}.get(n, b["fwd"])
        assert e < tol, (n, e)
    for k, e in r["loss"].items():
        assert e < b["loss"], (k, e)
    g = r["grad"]
    assert r["frozen_zero"]
    assert g["global"] < b["g_global"], g["global"]
    assert g["median"] < b["g_median"], g["median"]
    assert g["worst"] < b["g_worst"], (g["worst_key"], g["worst"])


@pytest.mark.parametrize("stage", [3, 4])
def test_fastpitch_five_step_trajectory(lib, stage):
    """Five optimizer steps on both sides from the same state: the loss of every step within 1e-3 of the oracle's."""
    traj, werr = pu.fastpitch_trajectory(stage, 4, 40, 150, True, steps=5)
    for t in traj:
        assert t["rel"] < 1e-4, t                                   # measured <= 3.6e-5 (stage 3), 2.2e-5 (stage 4)
    assert werr["global"] < 5e-5, werr["global"]                    # weights after 5 LAMB steps: measured 2.3e-5 / 1.1e-5
    assert werr["worst"] < 1.6e-3, (werr["worst_key"], werr["worst"])   # measured 7.7e-4 (a qkv bias)


def test_hifigan_step_at_baseline_shape(lib):
    """Measured at 16 x 8192 (profiles/r02_parity_table.txt): every loss term <= 1.7e-5; D-step gradients global 6.5e-4,
    worst tensor 2.2e-2 (a 15-tap first-layer weight_v of MSD scale 2); G-step gradients global 3.1e-3, worst 4.5e-3; weights
    after both AdamW steps: G 5.1e-4, MPD 2.8e-4, MSD 2.6e-4."""
    r = pu.hifigan_step_errors(16, 32, steps=1)
    for k, e in r["loss"][0].items():
        assert e < 4e-5, (k, e)
    assert r["dgrad"]["global"] < 1.3e-3, r["dgrad"]["global"]
    assert r["dgrad"]["worst"] < 4.4e-2, (r["dgrad"]["worst_key"], r["dgrad"]["worst"])
    assert r["ggrad"]["global"] < 6.2e-3, r["ggrad"]["global"]
    # worst of 234 tensors: 4.5e-3 (a 32-element weight_g) with the scalar weight-pack kernel, 1.0e-2 with the float4 one,
    # whose norm sums in another order and so rounds a few packed weights the other way -- tf32 noise, not a trend
    assert r["ggrad"]["worst"] < 2e-2, (r["ggrad"]["worst_key"], r["ggrad"]["worst"])
    for name, bound in (("G", 1.1e-3), ("mpd", 6e-4), ("msd", 6e-4)):
        assert r["weights"][name]["global"] < bound, (name, r["weights"][name]["global"])


def test_hifigan_three_step_trajectory(lib):
    """Three consecutive steps at 2 x 2048: measured loss differences <= 4.7e-5 / 3.0e-4 / 9.5e-4 (steps 0 / 1 / 2; the first
    AdamW steps are sign-like, so every step amplifies the tf32 rounding of near-zero gradient entries)."""
    r = pu.hifigan_step_errors(2, 8, steps=3)
    for s, l in enumerate(r["loss"]):
        for k, e in l.items():
            assert e < (1e-4, 6e-4, 2e-3)[s], (s, k, e)
