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
    a, b = a.detach(), b.detach()
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


HG_PATCHES = [('if dev.type != "cuda":', "if False:")]
VT_PATCHES = [('if dev.type != "cuda":', "if False:"),
              ('dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")', 'dev = torch.device("cpu")'),
              ('capi.call("xva_device_check", dev.index or 0)', "pass")]


def test_hifigan_generator_through_the_emulator_matches_the_oracle():
    """HiFi-GAN v1 Generator of the product package (hifigan/models.py:82-128: weight-norm packing of all 72 convolutions in
    one call, conv_pre, 4 x [2-phase transposed convolution + 3 ResBlock1 + MRF mean], conv_post + tanh) forward and its
    hand-written backward through the emulated C ABI vs oracle.hifigan.generator + autograd: output and all 234 parameter
    gradients (weight_g / weight_v through the weight-norm backward)."""
    from oracle import hifigan as ohg

    class H(dict):
        __getattr__ = dict.__getitem__

    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
    sd = ohg.make_generator_state(11, scale=0.7)
    gen = torch.Generator().manual_seed(11)
    mel = torch.randn(1, 80, 4, generator=gen)
    w = torch.randn(1, 1, 256 * 4, generator=gen)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ohg.generator(leaves, mel)
    (want * w).sum().backward()
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        G = hg.Generator(h, device="cpu")
        res = G.load_state_dict(sd)
        assert not res.missing_keys and not res.unexpected_keys
        G.train()
        y = G(mel)
        G.zero_grad()
        G.backward(w)
        grads = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
        used = set(cabi_emu.calls)
    assert {"xva_wn_pack_fwd", "xva_wn_pack_bwd", "xva_mean3_lrelu", "xva_sum3", "xva_tanh_bwd"} <= used
    assert y.shape == want.shape and rel(y, want.detach()) < 1e-5
    assert len(grads) == 234
    # (single leaky-ReLU sign decisions on pre-activations within fp32 rounding of zero can differ: bound per tensor 1e-3)
    worst = max((rel(grads[k], leaves[k].grad), k) for k in grads)
    assert worst[0] < 1e-3, worst


@pytest.mark.parametrize("with_g", [True, False])
def test_wavenet_stack_through_the_emulator_matches_the_oracle(with_g):
    """vits.WN (python/xvapitch/wavenet.py:16-106) stand-alone through the emulated C ABI: output, input and conditioning
    gradients, every parameter gradient vs autograd through oracle.vits.wn, ragged mask."""
    from oracle import vits as ov

    Hc, L, K, Cc = 64, 3, 5, 64
    gen = torch.Generator().manual_seed(9)
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        vt = cabi_emu.load_module("vits", VT_PATCHES, extra_modules={"xva_trainer_b200.hifigan": hg})
        m = vt.WN(Hc, Hc, K, 1, L, c_in_channels=Cc)
        sd = {}
        for k, p in m.named_parameters():
            if k.endswith("weight_v"):
                sd[k] = torch.randn(p.shape, generator=gen) * 0.7 / np.sqrt(p.shape[1] * p.shape[2])
            elif k.endswith("bias"):
                sd[k] = (torch.rand(p.shape, generator=gen) * 2 - 1) * 0.05
        for k, p in m.named_parameters():
            if k.endswith("weight_g"):
                sd[k] = sd[k[:-1] + "v"].flatten(1).norm(dim=1).view(p.shape) * (1 + 0.1 * torch.rand(p.shape, generator=gen))
        assert not m.load_state_dict(sd).missing_keys
        m.train()
        B, T = 2, 21
        x = torch.randn(B, Hc, T, generator=gen)
        g = torch.nn.functional.normalize(torch.randn(B, Cc, 1, generator=gen), dim=1) if with_g else None
        lens = [21, 13]
        mask = ov.sequence_mask(lens, T)[:, None, :].float()
        d = torch.randn(B, Hc, T, generator=gen)
        leaves = {f"enc.{k}": v.clone().requires_grad_(True) for k, v in sd.items()}
        xl = x.clone().requires_grad_(True)
        gl = g.clone().requires_grad_(True) if with_g else None
        want = ov.wn(leaves, "enc", xl, mask, gl, num_layers=L, hidden=Hc, kernel=K)
        (want * d).sum().backward()
        got = m(x, mask, g)
        m.zero_grad()
        dx, dg = m.backward(d)
        grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    assert rel(got, want.detach()) < 1e-5 and rel(dx, xl.grad) < 1e-5
    if with_g:
        assert rel(dg, gl.grad) < 1e-5
    for k, gr in grads.items():
        if k.startswith("cond_layer") and not with_g:
            continue
        assert rel(gr, leaves[f"enc.{k}"].grad) < 1e-4, k


def test_normalising_flow_through_the_emulator_matches_the_reference_recording():
    """vits.ResidualCouplingBlocks (python/xvapitch/model.py:1358-1421: 4 flows x [1x1, 4-layer WaveNet, 1x1, coupling, flip])
    through the emulated C ABI: forward and reverse outputs and the input / conditioning gradients as RECORDED from the
    reference module (tests/golden/vits_flow.npz), every parameter gradient vs the oracle's autograd and the recorded norms,
    reverse(forward(x)) = x."""
    from oracle import vits as ov
    from test_oracle_golden import _vits_flow_fixture

    gold, spec, sd = _vits_flow_fixture()
    x, cond, w = (torch.from_numpy(gold[k]) for k in ("x", "g", "w"))
    lens = [int(v) for v in gold["lens"]]
    mask = ov.sequence_mask(lens, x.shape[2])[:, None, :].float()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (ov.residual_coupling_blocks(leaves, x, mask, cond) * w).sum().backward()
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        vt = cabi_emu.load_module("vits", VT_PATCHES, extra_modules={"xva_trainer_b200.hifigan": hg})
        flow = vt.ResidualCouplingBlocks(192, 192, kernel_size=5, dilation_rate=1, num_layers=4, cond_channels=512, device="cpu")
        assert [(k, tuple(v.shape)) for k, v in flow.state_dict().items()] == [(k, tuple(sh)) for k, sh in spec]
        res = flow.load_state_dict(sd)
        assert not res.missing_keys and not res.unexpected_keys
        flow.train()
        z_p = flow(x, mask, g=cond)
        flow.zero_grad()
        dx, dg = flow.backward(w)
        grads = {k: p.grad.detach().clone() for k, p in flow.named_parameters()}
        with torch.no_grad():
            back = flow(z_p, mask, g=cond, reverse=True)
    assert rel(z_p, torch.from_numpy(gold["z_p"])) < 2e-5
    assert rel(dx, torch.from_numpy(gold["dx"])) < 5e-5 and rel(dg, torch.from_numpy(gold["dg"])) < 5e-5
    for (k, gr), want_norm in zip(grads.items(), gold["grad_norms"]):
        assert rel(gr, leaves[k].grad) < 2e-4, k
        assert abs(float(gr.double().norm()) - want_norm) < 2e-4 * want_norm + 1e-12, k
    assert rel(back, torch.from_numpy(gold["reverse_of_z_p"])) < 2e-5 and rel(back, x) < 2e-5


def test_hifigan_training_step_through_the_emulator_matches_the_oracle():
    """One HiFiTrainer.iteration body of the product package (hifigan.HiFiGANStep: generator forward, loss mel, D step over
    MPD + MSD incl. the spectral-normed scale discriminator and the grouped convolutions, AdamW; G step with 45 L1 mel +
    feature + adversarial losses, AdamW) through the emulated C ABI vs oracle.hifigan.train_step, which is pinned to two
    steps recorded from the unmodified reference (tests/test_oracle_golden.py): every loss term and all 404 updated tensors.
    The first AdamW step moves a weight by lr * sign(grad), so entries whose gradient is rounding noise may step the other
    way: the weights are compared at 2e-3 of the tensor norm, the losses at 1e-5."""
    from oracle import hifigan as ohg

    class H(dict):
        __getattr__ = dict.__getitem__

    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99,
          n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000, fmax_for_loss=None)
    sd_g = ohg.make_generator_state(5, scale=0.7)
    sd_p = ohg.make_disc_state(ohg.mpd_spec(), 21)
    sd_s = ohg.make_disc_state(ohg.msd_spec(), 22)
    x, y, y_mel = ohg.synthetic_batch(2, 8, seed=3)
    og, op, os_ = ({k: v.clone() for k, v in d.items()} for d in (sd_g, sd_p, sd_s))
    want, _ = ohg.train_step(og, op, os_, x, y, y_mel, {})
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        G = hg.Generator(h, device="cpu"); G.load_state_dict(sd_g); G.train()
        mpd = hg.MultiPeriodDiscriminator(device="cpu"); mpd.load_state_dict(sd_p); mpd.train()
        msd = hg.MultiScaleDiscriminator(device="cpu"); msd.load_state_dict(sd_s); msd.train()
        step = hg.HiFiGANStep(G, mpd, msd, h)
        losses = step.step(x, y, y_mel)
        after = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("G", G), ("mpd", mpd), ("msd", msd))}
        used = set(cabi_emu.calls)
    assert {"xva_sn_pack_fwd", "xva_sn_pack_bwd", "xva_conv_c1_fwd", "xva_conv_c1_bwd_x", "xva_avgpool4_fwd", "xva_spec_mag_bwd",
            "xva_l1_loss_grad", "xva_adamw_step"} <= used
    for k in ("loss_disc_all", "loss_mel", "loss_fm", "loss_gen", "loss_gen_all"):
        a, b = float(losses[k]), float(want[k])
        assert abs(a - b) < 1e-5 * abs(b) + 1e-7, (k, a, b)
    for name, ref, before in (("G", og, sd_g), ("mpd", op, sd_p), ("msd", os_, sd_s)):
        moved = 0
        for k, v in ref.items():
            assert rel(after[name][k], v) < 2e-3, (name, k, rel(after[name][k], v))
            moved += int(not torch.equal(after[name][k], before[k]))
        assert moved >= len(ref) * 0.9, (name, moved, len(ref))


def test_xvapitch_hifi_only_step_through_the_emulator_matches_the_reference_recording():
    """One xVAPitch --hifi_only iteration of the product package (vits.HifiOnlyStep: posterior encoder with its 16-layer
    WaveNet stack, segment draw, conditioned waveform decoder, VITS discriminator in one batched pass, TorchSTFT log-mels,
    both losses, both AdamW) through the emulated C ABI, the reference's random draws replayed, vs the iteration RECORDED
    from the unmodified reference modules (tests/golden/vits_hifi_only.npz): segment starts identical, every loss to 2e-5,
    the norm of every one of the 447 updated tensors and of its change; and vs the oracle step tensor by tensor."""
    from oracle import vits as ov
    from test_oracle_golden import _vits_hifi_only_fixture

    gold, specs, sds, linear, waveform, d_vectors = _vits_hifi_only_fixture()
    before = {n: {k: v.clone() for k, v in sd.items()} for n, sd in sds.items()}
    eps, u = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["u"])
    lens = [int(v) for v in gold["y_lengths"]]
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        vt = cabi_emu.load_module("vits", VT_PATCHES, extra_modules={"xva_trainer_b200.hifigan": hg})
        enc = vt.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device="cpu")
        dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2], inference_padding=0,
                                  cond_channels=512, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                                  conv_post_bias=False, device="cpu")
        disc = hg.VitsDiscriminator(device="cpu")
        for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
            assert [(k, tuple(v.shape)) for k, v in mod.state_dict().items()] == [(k, tuple(sh)) for k, sh in specs[name]], name
            res = mod.load_state_dict({k: v.clone() for k, v in sds[name].items()})
            assert not res.missing_keys and not res.unexpected_keys
            mod.train()
        losses = vt.HifiOnlyStep(enc, dec, disc).step(linear, lens, waveform, d_vectors, eps=eps, u=u)
        after = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("enc", enc), ("dec", dec), ("disc", disc))}
    assert losses["slice_ids"].tolist() == gold["slice_ids"].tolist()
    for k in ("loss", "loss_gen", "loss_feat", "loss_mel", "loss_disc"):
        a, b = float(losses[k]), float(gold[k])
        assert abs(a - b) < 2e-5 * abs(b), (k, a, b)
    osd = {n: {k: v.clone() for k, v in sd.items()} for n, sd in before.items()}
    ov.hifi_only_step(osd["enc"], osd["dec"], osd["disc"], linear, waveform, d_vectors, lens, eps, u, {})
    for name in ("enc", "dec", "disc"):
        moved = 0
        for i, (k, _) in enumerate(specs[name]):
            assert rel(after[name][k], osd[name][k]) < 2e-3, (name, k, rel(after[name][k], osd[name][k]))
            moved += int(not torch.equal(after[name][k], before[name][k]))
            n_after = float(after[name][k].double().norm())
            assert abs(n_after - gold[f"{name}/after_norms"][i]) < 1e-4 * gold[f"{name}/after_norms"][i] + 1e-9, (name, k)
            delta = float((after[name][k].double() - before[name][k].double()).norm())
            assert abs(delta - gold[f"{name}/delta_norms"][i]) < 0.05 * gold[f"{name}/delta_norms"][i] + 1e-9, (name, k, delta)
        assert moved == len(specs[name]), (name, moved)


def test_fastpitch_stage1_step_through_the_emulator_matches_the_oracle():
    """Training stage 1 (the aligner) of the product package through the emulated C ABI -- ConvAttention's projection stacks
    on the tap-GEMM, the score kernel, MAS, AttentionCTCLoss (here through torch's own CTC: an implementation independent
    of the kernel's recursion), AttentionBinarizationLoss, the combined gradient and the backward through both stacks --
    vs oracle.fastpitch: soft attention and log-probabilities, the hard alignment and durations, both losses, the gradient
    of the 11 stage-1 tensors, zero gradient everywhere else."""
    x, y = ofp.synthetic_batch(3, 14, 45, seed=21, ragged=True, prior=True)
    x[2] = x[2] * 4.0
    y[0] = x[2]
    sd = ofp.make_state(4321)
    klw = 0.5
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        m = fp.FastPitch(device="cpu")
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.training_stage = 1
        m.train()
        m.p_drop = 0.0
        crit = fp.FastPitchLoss()
        crit.training_stage = 1
        kl = fp.AttentionBinarizationLoss()
        out = m(x)
        loss, meta = crit(out, y)
        klv = kl(out[9], out[8])
        m.zero_grad()
        m.backward(crit, 1.0, kl=(kl, klw))
        got = m.grads()
        keys = fp.trainable_keys(1)
        used = set(cabi_emu.calls)
    assert {"xva_attn_score_fwd", "xva_attn_score_bwd", "xva_attn_ctc", "xva_mas_width1", "xva_attn_bin_loss",
            "xva_attn_grad_combine"} <= used
    want = ofp.forward(sd, x, 1)
    assert rel(out[8], want[8]) < 2e-5 and rel(out[11], want[11]) < 2e-5
    assert torch.equal(out[9], want[9]) and torch.equal(out[10], want[10])
    wtotal, wmeta = ofp.loss(want, y, 1, kl_weight=klw)
    assert abs(float(meta["attn_loss"]) - float(wmeta["attn_loss"])) <= 1e-5 * abs(float(wmeta["attn_loss"]))
    assert abs(klw * float(klv) - float(wmeta["kl_loss"])) <= 1e-5 * abs(float(wmeta["kl_loss"])) + 1e-9
    _, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, 1, 1e-3, {}, drop=0.0, training=False,
                               kl_weight=klw, clip=1e9)
    assert keys == ofp.trainable_keys(1) and len(keys) == 11
    errs = []
    for k, gr in got.items():
        if k not in keys:
            assert float(gr.abs().max()) == 0.0, k
        else:
            errs.append(rel(gr, wgrads[k]))
            assert errs[-1] < 2e-2, (k, errs[-1])      # (a handful of ReLU gates within fp32 rounding of zero, as on the device)
    assert sorted(errs)[len(errs) // 2] < 2e-5, sorted(errs)


def test_xvapitch_alignment_block_through_the_emulator():
    """vits.prior_alignment / prior_expand_backward / kl_loss (xvapitch/model.py:763-777, 855-856; losses.py:86-103) through
    the emulated C ABI vs oracle.vits: log-likelihood matrix, path and durations, expanded prior and its backward, KL loss
    and its four gradients."""
    from oracle import vits as ov

    gen = torch.Generator().manual_seed(4)
    B, C, tx, ty = 2, 64, 7, 23
    x_lens, y_lens = [7, 5], [23, 16]
    z_p = torch.randn(B, C, ty, generator=gen)
    m_p = torch.randn(B, C, tx, generator=gen)
    logs_p = torch.randn(B, C, tx, generator=gen) * 0.3
    logs_q = torch.randn(B, C, ty, generator=gen) * 0.3
    d_m, d_l = torch.randn(B, C, ty, generator=gen), torch.randn(B, C, ty, generator=gen)
    mp, lp = m_p.clone().requires_grad_(True), logs_p.clone().requires_grad_(True)
    w_logp, w_path, w_durs, w_me, w_le = ov.prior_alignment(z_p, mp, lp, x_lens, y_lens)
    ((w_me * d_m).sum() + (w_le * d_l).sum()).backward()
    zq, lq, me, le = (t.clone().requires_grad_(True) for t in (z_p, logs_q, w_me.detach(), w_le.detach()))
    z_mask = ov.sequence_mask(y_lens, ty)[:, None, :].float()
    w_kl = ov.kl_loss(zq, lq, me, le, z_mask)
    w_kl.backward()
    with cabi_emu.installed():
        hg = cabi_emu.load_module("hifigan", HG_PATCHES)
        vt = cabi_emu.load_module("vits", VT_PATCHES, extra_modules={"xva_trainer_b200.hifigan": hg})
        res = vt.prior_alignment(z_p, m_p, logs_p, x_lens, y_lens)
        dm, dl = vt.prior_expand_backward(d_m, d_l, res["cum"], tx)
        kl, grads = vt.kl_loss(z_p, logs_q, w_me.detach(), w_le.detach(), y_lens)
        used = set(cabi_emu.calls)
    assert {"xva_vits_logp_operands", "xva_mas_width1", "xva_regulate_len_fwd", "xva_regulate_len_bwd", "xva_vits_kl"} <= used
    valid = (ov.sequence_mask(x_lens, tx)[:, :, None] & ov.sequence_mask(y_lens, ty)[:, None, :])
    assert rel(res["logp"].transpose(1, 2) * valid, w_logp * valid) < 2e-6
    assert torch.equal(res["attn"][:, 0], w_path) and torch.equal(res["durations"][:, 0], w_durs)
    assert rel(res["m_p"], w_me) < 1e-6 and rel(res["logs_p"], w_le) < 1e-6
    assert rel(dm, mp.grad) < 1e-6 and rel(dl, lp.grad) < 1e-6
    assert abs(float(kl) - float(w_kl.detach())) < 1e-6 * abs(float(w_kl.detach()))
    for gr, w in zip(grads, (zq.grad, lq.grad, me.grad, le.grad)):
        assert rel(gr, w) < 1e-5


def test_every_compute_entry_point_of_the_abi_has_a_stand_in():
    """include/xva_b200.h via capi.PROTOTYPES: everything that launches work is emulated (or, for csrc/relattn.cu, compiled
    for the host); what is left are the library's introspection calls."""
    from xva_trainer_b200 import capi

    have = set(cabi_emu.TABLE) | set(cabi_emu.HOST_COMPILED)
    missing = sorted(set(capi.PROTOTYPES) - have)
    assert missing == ["xva_abi_version", "xva_attn_ctc_workspace_bytes", "xva_gemm_debug_counters", "xva_last_error",
                       "xva_sizeof_gemm_args", "xva_sizeof_sn_desc", "xva_sizeof_wn_desc"], missing


@pytest.mark.parametrize("case", ["free", "pace", "forced"])
def test_fastpitch_infer_through_the_emulator_matches_the_reference_recording(case):
    """FastPitch.infer of the product package (fastpitch/model.py:426-482) through the emulated C ABI vs the outputs
    RECORDED from the reference's own FastPitch.infer (tests/golden/infer.npz): predicted durations / pitch / energy, frame
    counts and the mel -- free-running, paced, and with forced durations + pitch."""
    from test_infer_gpu import GOLD, infer_kwargs, infer_state

    g = np.load(os.path.join(GOLD, "infer.npz"))
    t = lambda k: torch.from_numpy(g[k])
    kw = infer_kwargs(g, case)
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        m = fp.FastPitch(device="cpu")
        m.load_state_dict(infer_state())
        m.eval()
        mel, dec_lens, dur_pred, pitch_pred, energy_pred = m.infer(t("text"), **kw)
    assert rel(dur_pred, t(f"{case}/dur_pred")) < 2e-5
    assert rel(pitch_pred, t(f"{case}/pitch_pred")) < 2e-5 and rel(energy_pred, t(f"{case}/energy_pred")) < 2e-5
    assert torch.equal(dec_lens, t(f"{case}/dec_lens")) and dec_lens.dtype == torch.int64
    assert tuple(mel.shape) == tuple(g[f"{case}/mel"].shape) and rel(mel, t(f"{case}/mel")) < 2e-5


@pytest.mark.parametrize("stage", [3, 1])
def test_gradient_accumulation_over_micro_batches_is_the_sum_of_their_gradients(stage):
    """FastPitchTrainer accumulates `gam` micro-batches before one LAMB step (xva_train.py:806-813, 855-862: loss / gam,
    backward, step every gam-th): backward(criterion, 1 / gam) must ADD into the gradient arena -- every weight-gradient
    GEMM, column sum, LayerNorm / embedding / scalar-convolution gradient -- so that two micro-batches leave exactly
    (g_1 + g_2) / 2 behind."""
    xa, ya = ofp.synthetic_batch(2, 12, 40, seed=7, ragged=True, prior=(stage == 1))
    xb, yb = ofp.synthetic_batch(2, 12, 40, seed=8, ragged=True, prior=(stage == 1))
    sd = ofp.make_state(1234)
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)
        m = fp.FastPitch(device="cpu")
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.training_stage = stage
        m.train()
        m.p_drop = 0.0
        crit = fp.FastPitchLoss()
        crit.training_stage = stage
        kl = fp.AttentionBinarizationLoss()

        def micro(x, y, scale):
            out = m(x)
            crit(out, y)
            if stage == 1:
                kl(out[9], out[8])
                m.backward(crit, scale, kl=(kl, 0.5))          # (the binarization weight; backward multiplies it by `scale`)
            else:
                m.backward(crit, scale)

        singles = []
        for x, y in ((xa, ya), (xb, yb)):
            m.zero_grad()
            micro(x, y, 1.0)
            singles.append(m.arena.g.clone())
        m.zero_grad()
        micro(xa, ya, 0.5)
        micro(xb, yb, 0.5)
        both = m.arena.g.clone()
    want = 0.5 * (singles[0] + singles[1])
    assert float(want.norm()) > 0 and rel(both, want) < 1e-6
