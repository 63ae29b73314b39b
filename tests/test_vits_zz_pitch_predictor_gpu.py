"""xVAPitch pitch predictor (textenc.RelativePositioningPitchEnergyEncoder) on the device vs the recording of the unmodified
reference module (tests/golden/vits_pitch_predictor.npz) and the CPU oracle + autograd.

STATUS: written after the round's GPU budget was spent -- these tests have NOT run on hardware yet. What has run:
  * the module's host code on the CPU through the emulated C ABI against the same golden and oracle
    (tests/test_vits_text_encoder_cpu.py::test_pitch_predictor_*), the method that predicted 12 / 12 for the text encoder
    (profiles/r02_textenc_gpu_tests.log);
  * on the B200: every kernel it launches -- the transformer layers are the text encoder's (tests/test_vits_text_encoder_gpu.py),
    LayerNorm at 708 / 780 channels (tests/test_rowops_gpu.py, call AT), the N = 1 / M = 1 tap-GEMM shapes of the
    projection (HiFi-GAN conv_post, tests/test_hifigan_gpu.py).
They are therefore marked xfail(strict=False): an unexpected failure on the first hardware run is reported (xfailed) without
hiding the state of the rest of the suite, a pass shows up as xpassed. Remove the mark after the first green run."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from textenc_util import fill_pitch, pitch_ref_spec, rel  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first hardware run pending: written after the round's GPU budget was spent; "
                                                     "host code verified on the CPU through the emulated C ABI")]


def _case(layers, hidden, T, lens, seed):
    from oracle import vits as ov

    gen = torch.Generator().manual_seed(seed)
    sd = fill_pitch(pitch_ref_spec(layers=layers, hidden=hidden), gen)
    B = len(lens)
    x = torch.randn(B, T, hidden, generator=gen)
    spk = torch.nn.functional.normalize(torch.randn(B, 512, 1, generator=gen), dim=1)
    r = torch.randn(B, 1, T, generator=gen)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ov.pitch_predictor(p, x, lens, spk, num_layers=layers)
    (want * r).sum().backward()
    return sd, x, spk, r, want.detach(), {k: v.grad for k, v in p.items() if v.grad is not None}


def _run(sd, layers, hidden, x, lens, spk, r):
    from xva_trainer_b200 import textenc

    m = textenc.RelativePositioningPitchEnergyEncoder(1, hidden, 768, 2, layers, 3, 0.0, conditioning_emb_dim=512)
    m.load_state_dict(sd)
    m.train()
    m.zero_grad()
    pred = m(x.cuda(), lens, speaker_emb=spk.cuda())
    m.backward(r.cuda())
    torch.cuda.synchronize()
    return m, pred, m.grads()


CASES = [(3, 196, 13, [13, 8], 71), (2, 268, 40, [40, 17, 33], 5)]


@pytest.mark.parametrize("layers,hidden,T,lens,seed", CASES)
def test_wiring_exact_with_fp32_checker_gemm(lib, monkeypatch, layers, hidden, T, lens, seed):
    from xva_trainer_b200 import capi, ops

    sd, x, spk, r, want, wgrads = _case(layers, hidden, T, lens, seed)
    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        m, pred, got = _run(sd, layers, hidden, x, lens, spk, r)
    finally:
        capi.call("xva_set_operand_rounding", 1)
    assert rel(pred, want) < 2e-5
    assert set(got) == set(wgrads)
    floor = 1e-4 * max(float(v.norm()) for v in wgrads.values())
    for k, w in wgrads.items():
        assert float((got[k].cpu() - w).norm()) / max(float(w.norm()), floor) < 2e-4, k


@pytest.mark.parametrize("layers,hidden,T,lens,seed", CASES)
def test_product_path_matches_the_oracle(lib, layers, hidden, T, lens, seed):
    """Bounds = those of the text encoder's product path (measured there: forward 6e-4, gradient vector 9e-3, worst tensor
    2.9e-2, profiles/r02_textenc.txt); to be replaced by 2 x this module's own measurements after its first run."""
    sd, x, spk, r, want, wgrads = _case(layers, hidden, T, lens, seed)
    m, pred, got = _run(sd, layers, hidden, x, lens, spk, r)
    assert rel(pred, want) < 3e-3
    for b, n in enumerate(lens):
        if n < T:
            assert float(pred[b, :, n:].abs().max()) == 0.0
    num = sum(float((got[k].cpu() - w).norm()) ** 2 for k, w in wgrads.items())
    den = sum(float(w.norm()) ** 2 for w in wgrads.values())
    assert math.sqrt(num / den) < 2e-2
    floor = 1e-2 * max(float(v.norm()) for v in wgrads.values())
    for k, w in wgrads.items():
        assert float((got[k].cpu() - w).norm()) / max(float(w.norm()), floor) < 6e-2, k


def test_forward_matches_the_reference_golden_and_dead_state_survives_adamw(lib):
    from xva_trainer_b200 import hifigan, textenc

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vits_pitch_predictor.npz"))
    gen = torch.Generator().manual_seed(71)
    sd = fill_pitch(pitch_ref_spec(), gen)
    x, spk, r = torch.from_numpy(g["x"]), torch.from_numpy(g["spk"]), torch.from_numpy(g["r"])
    m = textenc.RelativePositioningPitchEnergyEncoder(1, 196, 768, 2, 3, 3, 0.1, conditioning_emb_dim=512)
    m.load_state_dict(sd)
    m.eval()
    pred = m(x.cuda(), [13, 8], speaker_emb=spk.cuda())
    assert rel(pred, torch.from_numpy(g["pitch_pred"])) < 3e-3
    opt = hifigan.AdamW([m.flat], lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
    opt.zero_grad()
    m.train()
    m(x.cuda(), [13, 8], speaker_emb=spk.cuda())
    m.backward(r.cuda())
    opt.step()
    torch.cuda.synchronize()
    after = m.state_dict()
    assert list(after) == list(sd)
    for k in m.dead_keys():                                   # the reference never gives these a gradient: AdamW skips them
        assert torch.equal(after[k].cpu(), sd[k]), k
    moved = [k for k in sd if k not in m.dead_keys() and not torch.equal(after[k].cpu(), sd[k])]
    assert len(moved) == len(sd) - 6
