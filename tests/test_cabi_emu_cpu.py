"""Validates tests/cabi_emu.py (the CPU stand-in for part of the C ABI) on host code whose GPU parity is established: one
FFT block of FastPitch (fastpitch.FastPitch._layer_fwd / _layer_bwd through the un-fused attention chain -- tap-GEMM modes
0 / 1 / 2 with 1 and 3 taps, bias, ReLU, gate, residual, length masking, softmax forward / backward, LayerNorm forward /
backward, column sums) against oracle.fastpitch + autograd. If the emulator follows include/xva_b200.h, the result must
match the fp32 oracle to rounding; the new host code of the xVAPitch text encoder is then checked the same way
(tests/test_vits_text_encoder_cpu.py). CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cabi_emu  # noqa: E402
from oracle import fastpitch as ofp  # noqa: E402

FP_PATCHES = [('if self.device_.type != "cuda":', "if False:"),
              ('self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
               'self.device_ = torch.device("cpu")')]


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("stack,T,lens", [("decoder", 37, [37, 29, 11]), ("encoder", 16, [16, 16, 9])])
def test_fft_block_through_the_emulator_matches_the_oracle(stack, T, lens):
    torch.manual_seed(3)
    sd = ofp.make_state(seed=11, perturb=True)
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        m = fp.FastPitch(device="cpu")
        m.load_state_dict(sd)
        m.eval()
        m.fused_attn = False
        B, D = len(lens), fp.D_MODEL
        li = torch.tensor(lens, dtype=torch.int32)
        mask = (torch.arange(T)[None, :] < li[:, None]).unsqueeze(2)
        x = (torch.randn(B, T, D) * mask).contiguous()
        dy = (torch.randn(B, T, D) * mask).contiguous()
        L = (m.dec_layers if stack == "decoder" else m.enc_layers)[2]
        save = []
        m.zero_grad()
        y = m._layer_fwd(x, li, L, save)
        dx = m._layer_bwd(dy, li, L, save[0])
        got_grads = m.grads()
        used = set(cabi_emu.calls)
    assert {"xva_gemm", "xva_softmax_fwd", "xva_softmax_bwd", "xva_layernorm_fwd", "xva_layernorm_bwd", "xva_colsum"} <= used

    p = f"{stack}.layers.2"
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(p)}
    xr = x.clone().requires_grad_(True)
    want = ofp.multi_head_attn(xr, ~mask.squeeze(2), params, f"{p}.dec_attn") * mask
    want = ofp.conv_ff(want, params, f"{p}.pos_ff") * mask
    want.backward(dy)
    assert rel(y, want.detach()) < 2e-6
    assert rel(dx, xr.grad) < 2e-5
    for k, v in params.items():
        assert rel(got_grads[k], v.grad) < 3e-5, k


def test_dropout_hash_is_the_librarys():
    """csrc/common.cuh hash_u64 / dropout_scale restated in numpy: a few values computed by hand from the C definition,
    the keep rate, and the one-hash-per-four-elements layout."""
    idx = np.arange(4096, dtype=np.uint64)
    s = cabi_emu.dropout_scale(12345, None, idx, 0.25)
    assert set(np.unique(s).tolist()) <= {0.0, np.float32(1.0 / 0.75).item()}
    assert abs(float((s > 0).mean()) - 0.75) < 0.03

    def h64(seed, i4):
        M = (1 << 64) - 1
        x = (i4 * 0x9E3779B97F4A7C15 + seed) & M
        x ^= x >> 32
        x = (x * 0xD6E8FEB86659FD93) & M
        x ^= x >> 32
        x = (x * 0xD6E8FEB86659FD93) & M
        x ^= x >> 32
        return x

    th = int(0.25 * 4294967296.0) >> 16
    for i in (0, 1, 2, 3, 4, 777, 4095):
        f = (h64(12345, i >> 2) >> (16 * (i & 3))) & 0xFFFF
        assert (s[i] > 0) == (f >= th)
    assert (cabi_emu.dropout_scale(1, None, idx, 0.0) == 1.0).all()


@pytest.mark.parametrize("stage,fused", [(3, True), (2, False), (4, True)])
def test_fastpitch_training_steps_through_the_emulator_match_the_oracle(stage, fused):
    """The whole FastPitch micro-step of the product package -- forward (embedding, FFT stacks, predictors, pitch / energy
    embeddings, length regulator, projection), FastPitchLoss, the hand-written backward, clip + LAMB -- executed on the CPU
    through the emulated C ABI, two consecutive optimizer steps, against oracle.fastpitch.train_step: forward tensors,
    every loss term, every parameter gradient, the weights after LAMB. (On the device the same comparison is
    tests/test_fastpitch_gpu.py; here it pins the HOST code -- 330 launches per step and their arguments -- between GPU runs,
    numerically rather than as a fingerprint.) fused = the fused-attention entry points (emulated from their contract in
    include/xva_b200.h) or the six-launch chain."""
    x, y = ofp.synthetic_batch(3, 12, 40, seed=7, ragged=True)
    sd = ofp.make_state(1234)
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        m = fp.FastPitch(device="cpu")
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.training_stage = stage
        m.train()
        m.p_drop = 0.0
        m.fused_attn = fused
        crit = fp.FastPitchLoss()
        crit.training_stage = stage
        opt = fp.Lamb(m, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        osd, ostate = {k: v.clone() for k, v in sd.items()}, {}
        keys = fp.trainable_keys(stage)
        for it in (50000, 50001):
            fp.adjust_learning_rate(it, opt, 0.1, 1000)
            m.zero_grad()
            out = m(x)
            loss, meta = crit(out, y)
            m.backward(crit, 1.0)
            got_grads = {k: v.clone() for k, v in m.grads(keys).items()}
            want_fwd = ofp.forward(osd, x, stage)
            wmeta, wgrads = ofp.train_step(osd, x, y, stage, ofp.noam_lr(it), ostate, drop=0.0, training=False)
            for name, g_, w_ in zip(("mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred",
                                     "energy_tgt"), out[:8], want_fwd[:8]):
                if w_ is None:
                    continue
                if w_.dtype == torch.bool:
                    assert torch.equal(g_, w_), name
                else:
                    assert rel(g_.float(), w_.float()) < 2e-5, (it, name)
            for k in ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"):
                a, b = float(meta[k]), float(wmeta[k])
                assert abs(a - b) <= 2e-5 * max(abs(b), 1e-6), (it, k, a, b)
            floor = 1e-4 * max(float(w.norm()) for w in wgrads.values() if w is not None)
            for k in keys:
                if wgrads[k] is None:
                    assert float(got_grads[k].abs().max()) == 0.0, k
                else:
                    err = float((got_grads[k] - wgrads[k]).norm()) / max(float(wgrads[k].norm()), floor)
                    assert err < 2e-4, (it, k, err)
            opt.step()
            m.step_dropout()
            after = m.state_dict()
            for k in keys:
                assert rel(after[k], osd[k]) < 5e-5, (it, k)      # (LAMB divides by sqrt(v): tensors with ~0 gradients amplify rounding)
        used = set(cabi_emu.calls)
    want_used = {"xva_embed_pos", "xva_lamb_step", "xva_grad_sqnorm"} | (
        {"xva_lens_mse", "xva_rowdot_bwd"} if stage == 2 else {"xva_regulate_len_fwd", "xva_regulate_len_bwd", "xva_mel_mse"})
    assert want_used <= used, want_used - used
    assert ("xva_attn_fwd" in used) == fused
