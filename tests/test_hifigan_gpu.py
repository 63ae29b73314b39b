"""HiFi-GAN generator (xva-trainer_b200/hifigan.py, all math through libxva_b200.so) vs the CPU oracle (pinned to the
reference by tests/test_oracle_golden.py) and vs the golden fixture recorded from the reference's Generator.

Tolerances: tf32 tensor-core operands (rounded to nearest), fp32 accumulation. Forward waveform: relative L2 <= 2e-3
(49 convolutions deep). Parameter gradients: <= 1e-1 per tensor (the small bias
gradients of the deep stages), <= 4e-2 on the global vector (measured 1.6e-2 .. 2.7e-2 over four seeds,
scripts/diag_gen_grad.py: spread evenly over the layers and largest for ups.0 / conv_pre, whose backward
signal has passed through 98 tf32 products and as many leaky-ReLU gates); with the exact-fp32
checker GEMM and operand rounding off (wiring check): forward <= 1e-5, gradients <= 1e-4 (leaky ReLU has no dead zone, so
there is no gate-flip noise floor here as there is for FastPitch's ReLU)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hifigan as ohg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


class _H(dict):
    __getattr__ = dict.__getitem__


def _config():
    h = _H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
           resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
    return h


def _generator(lib, sd):
    from xva_trainer_b200 import hifigan as hg

    g = hg.Generator(_config(), device="cuda:0")
    missing = g.load_state_dict({k: v for k, v in sd.items()})
    assert not missing.missing_keys and not missing.unexpected_keys
    g.train()
    return g


def test_state_dict_keys_match_reference(lib):
    ref = json.load(open(os.path.join(GOLD, "hifigan_keys.json")))
    g = _generator(lib, ohg.make_generator_state(1))
    assert [[k, list(v.shape)] for k, v in g.state_dict().items()] == ref


def _run(g, sd, mel, w):
    y = g(mel.cuda())
    g.zero_grad()
    g.backward(w.cuda())
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu() for k, p in g.named_parameters()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    yo = ohg.generator(leaves, mel)
    (yo * w).sum().backward()
    return y, yo.detach(), grads, {k: p.grad for k, p in leaves.items()}


@pytest.mark.parametrize("T,seed", [(6, 3), (9, 4)])
def test_generator_matches_oracle(lib, T, seed):
    sd = ohg.make_generator_state(seed, scale=0.7)
    g = _generator(lib, sd)
    gen = torch.Generator().manual_seed(seed)
    mel = torch.randn(2, 80, T, generator=gen)
    w = torch.randn(2, 1, 256 * T, generator=gen)
    y, yo, grads, want = _run(g, sd, mel, w)
    assert y.shape == yo.shape
    assert rel(y, yo) < 2e-3, rel(y, yo)
    num = den = 0.0
    for k, gr in grads.items():
        e = rel(gr, want[k])
        assert e < 1e-1, (k, e)
        num += float((gr.double() - want[k].double()).pow(2).sum())
        den += float(want[k].double().pow(2).sum())
    assert (num / den) ** 0.5 < 4e-2, (num / den) ** 0.5


def test_generator_wiring_exact(lib, monkeypatch):
    from xva_trainer_b200 import capi, ops

    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        sd = ohg.make_generator_state(11, scale=0.7)
        g = _generator(lib, sd)
        gen = torch.Generator().manual_seed(11)
        mel = torch.randn(1, 80, 5, generator=gen)
        w = torch.randn(1, 1, 256 * 5, generator=gen)
        y, yo, grads, want = _run(g, sd, mel, w)
        assert rel(y, yo) < 1e-5, rel(y, yo)
        for k, gr in grads.items():
            assert rel(gr, want[k]) < 1e-4, (k, rel(gr, want[k]))
    finally:
        capi.call("xva_set_operand_rounding", 1)


def test_generator_matches_reference_golden(lib):
    gold = np.load(os.path.join(GOLD, "hifigan_small.npz"))
    sd = ohg.make_generator_state(1234, scale=1.0)
    g = _generator(lib, sd)
    y = g(torch.from_numpy(gold["gen/mel"]).cuda())
    assert rel(y, torch.from_numpy(gold["gen/y"])) < 2e-3
    g.zero_grad()
    g.backward(torch.from_numpy(gold["gen/w"]).cuda())
    for k, p in g.named_parameters():
        want_norm = float(gold[f"gen/grad/{k}/norm"])
        assert abs(float(p.grad.double().norm()) - want_norm) <= 3e-2 * want_norm + 1e-12, (k, float(p.grad.norm()), want_norm)


def _check_packer(pk, ref, params):
    """pk: _WnPacker; ref: {key: tuple of torch-packed tensors under autograd}. Forward within one tf32 rounding of the
    PyTorch weight-norm + re-layout; backward (gradient arena -> weight_g / weight_v) equal to autograd's to 1e-5."""
    W = pk.pack()
    for k, ts in ref.items():
        for a, b in zip(W[k], ts):
            assert a.shape == b.shape, (k, a.shape, b.shape)
            assert rel(a, b) < 5e-4, (k, rel(a, b))
            assert float((a[b == 0]).abs().max() if bool((b == 0).any()) else 0.0) == 0.0    # block-diagonal zeros
    gW = pk.zero_grads()
    gen = torch.Generator(device="cuda").manual_seed(5)
    tensors, grads = [], []
    for k, ts in ref.items():
        for g_, t in zip(gW[k], ts):
            g_.copy_(torch.randn(g_.shape, device="cuda", generator=gen))
            tensors.append(t)
            grads.append(g_.clone())
    for p_ in params:
        p_.grad = None
    torch.autograd.backward(tensors, grads)
    want = [p_.grad.clone() if p_.grad is not None else None for p_ in params]
    for p_ in params:
        p_.grad = None
    pk.unpack_grads()
    torch.cuda.synchronize()
    for p_, w in zip(params, want):
        if w is None:
            continue
        assert rel(p_.grad, w) < 1e-5, (tuple(p_.shape), rel(p_.grad, w))


def test_wn_packer_matches_torch_weight_norm(lib):
    """xva_wn_pack_fwd / _bwd (one launch for all convolutions) vs the PyTorch ops they replace."""
    from xva_trainer_b200 import hifigan as hg

    g = _generator(lib, ohg.make_generator_state(7, scale=0.7))
    _check_packer(g._get_packer(), g._pack(), [p for n, p in g.named_parameters() if "weight_" in n])
    for name in ("mpd", "msd"):
        _, m, _ = _disc_models(lib, name)
        pk = hg._WnPacker()
        ref, params = {}, []
        for i, d in enumerate(m.discriminators):
            if any(c.spectral for c in d.convs):
                continue
            d.register_weights(pk, str(i))
            for li, t in enumerate(d._packed()):
                ref[f"{i}.{li}"] = (t,)
            params += [p for n, p in d.named_parameters() if "weight_" in n]
        pk.finalize("cuda")
        _check_packer(pk, ref, params)


# ------------------------------------------------------------------------------------------------ mel spectrogram
@pytest.mark.parametrize("fmax", [8000, None])
def test_mel_spectrogram_matches_oracle_and_golden(lib, fmax):
    """Forward vs the reference fixture and the oracle; backward vs torch autograd through the oracle. The log-mel is
    compared in absolute terms (values span about [-11.5, 2]); tolerance 5e-3 absolute = tf32 DFT of a signal whose
    spectrum spans several decades, then log."""
    from xva_trainer_b200 import hifigan as hg

    gold = np.load(os.path.join(GOLD, "hifigan_small.npz"))
    audio = torch.from_numpy(gold["mel/audio"])
    ms = hg.MelSpectrogram(fmax=fmax, device="cuda:0")
    got = ms(audio.cuda()).transpose(1, 2).cpu()
    want = torch.from_numpy(gold[f"mel/{'none' if fmax is None else '8000'}"])
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 5e-3, float((got - want).abs().max())
    assert rel(got, want) < 1e-3
    # gradient of a random linear functional of the mel
    g = torch.Generator().manual_seed(3)
    w = torch.randn(want.shape, generator=g)
    a = audio.clone().requires_grad_(True)
    (ohg.mel_spectrogram(a, fmax) * w).sum().backward()
    d = ms.backward(w.transpose(1, 2).contiguous().cuda()).cpu()
    assert rel(d, a.grad) < 1e-2, rel(d, a.grad)


def test_tacotron_stft_variant_matches_reference_golden(lib):
    """TacotronSTFT.mel_spectrogram (reflect pad n_fft / 2, no epsilon, N / 256 + 1 frames) vs the
    fixture recorded from the reference and vs the oracle. Same tolerance as the HiFi-GAN variant (5e-3 absolute on the
    log-mel; tf32 DFT)."""
    from xva_trainer_b200 import hifigan as hg

    gold = np.load(os.path.join(GOLD, "tacotron_stft.npz"))
    audio = torch.from_numpy(gold["audio"])
    ms = hg.MelSpectrogram.tacotron(device="cuda:0")
    got = ms(audio.cuda()).transpose(1, 2).cpu()
    want = torch.from_numpy(gold["mel"])
    assert got.shape == want.shape, (got.shape, want.shape)
    assert float((got - want).abs().max()) < 5e-3, float((got - want).abs().max())
    assert rel(got, ohg.tacotron_mel(audio)) < 1e-3


def test_reflect_pad_and_losses(lib):
    from xva_trainer_b200 import ops

    g = torch.Generator().manual_seed(0)
    y = torch.randn(3, 700, generator=g).cuda()
    yp = ops.reflect_pad(y, 384)
    want = torch.nn.functional.pad(y.unsqueeze(1), (384, 384), mode="reflect").squeeze(1)
    assert rel(yp, want) < 3e-4          # tf32-rounded copy
    d = torch.randn(3, 700 + 768, generator=g).cuda()
    yy = y.clone().requires_grad_(True)
    (torch.nn.functional.pad(yy.unsqueeze(1), (384, 384), mode="reflect").squeeze(1) * d).sum().backward()
    assert rel(ops.reflect_pad_bwd(d, 700, 384), yy.grad) < 1e-6
    a, b = torch.randn(1000, generator=g).cuda(), torch.randn(1000, generator=g).cuda()
    acc = torch.zeros(2, device="cuda", dtype=torch.float64)
    ops.reduce_l1(a, b, acc[0:1])
    ops.reduce_sq(a, 1.0, acc[1:2])
    assert abs(float(acc[0]) - float((a - b).abs().sum())) < 1e-3
    assert abs(float(acc[1]) - float(((1 - a) ** 2).sum())) < 1e-3
    assert torch.equal(ops.l1_grad(a, b, 0.5), 0.5 * torch.sign(b - a))
    assert rel(ops.sq_grad(a, 1.0, 0.25), 0.5 * (a - 1.0)) < 1e-6


# ------------------------------------------------------------------------------------------------ discriminators
def _disc_models(lib, name):
    from xva_trainer_b200 import hifigan as hg

    spec, seed = (ohg.mpd_spec(), 21) if name == "mpd" else (ohg.msd_spec(), 22)
    sd = ohg.make_disc_state(spec, seed)
    m = (hg.MultiPeriodDiscriminator if name == "mpd" else hg.MultiScaleDiscriminator)(device="cuda:0")
    res = m.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    m.train()
    return hg, m, sd


def _to_ref_layout(f, B):
    """engine fmap [B*P, Lp, C] (channels-last, P period columns) -> reference [B, C, L, P] (or [B, C, L] when P = 1)"""
    Z, L, Cc = f.shape
    P = Z // B
    t = f.view(B, P, L, Cc).permute(0, 3, 2, 1)
    return t if P > 1 else t[..., 0]


@pytest.mark.parametrize("name", ["mpd", "msd"])
def test_discriminator_forward_matches_oracle_and_golden(lib, name):
    hg, m, sd = _disc_models(lib, name)
    gold = np.load(os.path.join(GOLD, "hifigan_small.npz"))
    y, yh = torch.from_numpy(gold["disc/y"]), torch.from_numpy(gold["disc/y_hat"])
    rs, gs, frs, fgs = m(y.cuda(), yh.cuda())
    ors, ogs, ofrs, ofgs = (ohg.mpd if name == "mpd" else ohg.msd)(sd, y, yh, training=True)
    B = y.shape[0]
    for i in range(len(rs)):
        want_r = torch.from_numpy(gold[f"disc/{name}/r{i}"])
        got_r = _to_ref_layout(rs[i], B).reshape(B, -1).cpu()
        assert rel(got_r, want_r) < 3e-3, (i, rel(got_r, want_r))
        assert rel(_to_ref_layout(gs[i], B).reshape(B, -1).cpu(), ogs[i]) < 3e-3
        for l, (f, of) in enumerate(zip(fgs[i], ofgs[i])):
            got = _to_ref_layout(f, B)[:, :, :of.shape[2]].cpu()
            assert got.shape == of.shape, (i, l, got.shape, of.shape)
            assert rel(got, of) < 3e-3, (i, l, rel(got, of))
            assert float(f[:, of.shape[2]:].abs().max() if f.shape[1] > of.shape[2] else 0.0) == 0.0   # alignment rows stay zero


@pytest.mark.parametrize("name", ["mpd", "msd"])
def test_discriminator_step_gradients(lib, name):
    """D step: discriminator_loss and every parameter gradient; G step: generator_loss + feature_loss and the gradient
    wrt the generated waveform. Oracle = torch autograd through oracle/hifigan.py."""
    hg, m, sd = _disc_models(lib, name)
    g = torch.Generator().manual_seed(31)
    B, T = 2, 1536
    y = 0.9 * torch.tanh(torch.randn(B, 1, T, generator=g))
    yh = 0.9 * torch.tanh(torch.randn(B, 1, T, generator=g))
    fn = ohg.mpd if name == "mpd" else ohg.msd
    # ---- D step
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("weight_u") and not (k.endswith("weight_v") and v.dim() == 1) else v.clone())
              for k, v in sd.items()}
    ors, ogs, _, _ = fn(leaves, y, yh, training=True)
    want_loss = ohg.discriminator_loss(ors, ogs)
    want_loss.backward()
    m.zero_grad()
    rs, gs, frs, fgs = m(y.cuda(), yh.cuda())
    loss = hg.discriminator_loss_backward(m, rs, gs)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(want_loss)) < 2e-3 * abs(float(want_loss))
    num = den = 0.0
    for k, p in m.named_parameters():
        w = leaves[k].grad
        e = rel(p.grad, w)
        num += float((p.grad.cpu().double() - w.double()).pow(2).sum())
        den += float(w.double().pow(2).sum())
        assert e < 6e-2, (k, e)
    assert (num / den) ** 0.5 < 2e-2
    # ---- G step (fresh spectral-norm state in both: reload)
    hg, m, sd = _disc_models(lib, name)
    yh_leaf = yh.clone().requires_grad_(True)
    ors, ogs, ofrs, ofgs = fn(sd, y, yh_leaf, training=True)
    want = ohg.generator_loss(ogs) + ohg.feature_loss(ofrs, ofgs)
    want.backward()
    rs, gs, frs, fgs = m(y.cuda(), yh.cuda())
    dwave = torch.zeros(B, T, device="cuda")
    lg, lf = hg.generator_adv_loss_backward(m, gs, frs, fgs, dwave, pools=(name == "msd"))
    torch.cuda.synchronize()
    assert abs(float(lg + lf) - float(want)) < 2e-3 * abs(float(want)), (float(lg), float(lf), float(want))
    assert rel(dwave.cpu(), yh_leaf.grad.reshape(B, T)) < 3e-2, rel(dwave.cpu(), yh_leaf.grad.reshape(B, T))


# ------------------------------------------------------------------------------------------------ full training step
def test_full_hifigan_step_matches_oracle(lib):
    """One HiFiTrainer.iteration (D step + G step, both AdamW updates) vs the oracle: every loss term within 2e-3 and
    the updated generator / discriminator weights within 5e-3 relative of the oracle's (the first AdamW step moves each
    weight by lr = 2e-4 in the direction of sign(grad) -- about 2 % of a weight here -- so this bounds the fraction of
    near-zero gradient entries whose sign differs under tf32 to well below 1 %)."""
    from xva_trainer_b200 import hifigan as hg

    h = _config()
    h.update(learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
             win_size=1024, fmin=0, fmax=8000, fmax_for_loss=None)
    sd_g = ohg.make_generator_state(5, scale=0.7)
    sd_p = ohg.make_disc_state(ohg.mpd_spec(), 21)
    sd_s = ohg.make_disc_state(ohg.msd_spec(), 22)
    G = _generator(lib, sd_g)
    mpd = hg.MultiPeriodDiscriminator(device="cuda:0"); mpd.load_state_dict(sd_p); mpd.train()
    msd = hg.MultiScaleDiscriminator(device="cuda:0"); msd.load_state_dict(sd_s); msd.train()
    step = hg.HiFiGANStep(G, mpd, msd, h)
    x, y, y_mel = ohg.synthetic_batch(2, 8, seed=3)
    losses = step.step(x.cuda(), y.cuda(), y_mel.cuda())
    torch.cuda.synchronize()
    og, op, os_ = ({k: v.clone() for k, v in d.items()} for d in (sd_g, sd_p, sd_s))
    want, _ = ohg.train_step(og, op, os_, x, y, y_mel, {})
    for k in ("loss_disc_all", "loss_mel", "loss_fm", "loss_gen", "loss_gen_all"):
        a, b = float(losses[k]), float(want[k])
        assert abs(a - b) < 2e-3 * abs(b) + 1e-6, (k, a, b)
    for name, model, ref, before in (("G", G, og, sd_g), ("mpd", mpd, op, sd_p), ("msd", msd, os_, sd_s)):
        after = model.state_dict()
        moved = 0
        for k, v in ref.items():
            assert rel(after[k], v) < 5e-3, (name, k, rel(after[k], v))
            moved += int(not torch.equal(after[k].cpu(), before[k]))
        assert moved >= len(ref) * 0.9, (name, moved, len(ref))


@pytest.mark.parametrize("mode", ["side_stream", "disc_branches", "gen_branches", "all"])
def test_two_stream_backward_gives_the_same_step(lib, mode):
    """hifigan._Side (XVA_BWD_STREAMS): weight / bias gradients on a side stream; hifigan._Branches (XVA_DISC_STREAMS /
    XVA_GEN_STREAMS): sub-discriminators / the ResBlocks of an MRF stage on parallel streams -- all three are the default
    inside a captured graph. The same two training steps from the same state must produce the same losses and the same
    updated weights up to the order of the fp32 atomic additions of the split weight gradients, which the streams
    change: measured 1.1e-5 on a loss of the second step (B200, profiles/r02_streams_ab.txt), bound 1e-4."""
    from xva_trainer_b200 import hifigan as hg

    h = _config()
    h.update(learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256,
             win_size=1024, fmin=0, fmax=8000, fmax_for_loss=None)
    x, y, y_mel = (t.cuda() for t in ohg.synthetic_batch(2, 8, seed=3))
    results = []
    was = (hg._Side.enabled, hg._Branches.n, hg._Branches.n_gen)
    try:
        for flag in (False, True):
            hg._Side.enabled = bool(flag and mode in ("side_stream", "all"))
            hg._Branches.n = 4 if (flag and mode in ("disc_branches", "all")) else 0
            hg._Branches.n_gen = 3 if (flag and mode in ("gen_branches", "all")) else 0
            G = _generator(lib, ohg.make_generator_state(5, scale=0.7))
            mpd = hg.MultiPeriodDiscriminator(device="cuda:0"); mpd.load_state_dict(ohg.make_disc_state(ohg.mpd_spec(), 21)); mpd.train()
            msd = hg.MultiScaleDiscriminator(device="cuda:0"); msd.load_state_dict(ohg.make_disc_state(ohg.msd_spec(), 22)); msd.train()
            step = hg.HiFiGANStep(G, mpd, msd, h)
            for _ in range(2):
                losses = step.step(x, y, y_mel)
            torch.cuda.synchronize()
            results.append(({k: float(v) for k, v in losses.items()},
                            {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("G", G), ("mpd", mpd), ("msd", msd))}))
    finally:
        hg._Side.enabled, hg._Branches.n, hg._Branches.n_gen = was
    (l0, s0), (l1, s1) = results
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 1e-4 * abs(l0[k]) + 1e-7, (k, l0[k], l1[k])
    # weights after two AdamW steps: an entry whose gradient is within atomic-order rounding of zero moves by +-lr in
    # either run (the first AdamW steps are sign-like), so the bound is the one the oracle comparison above uses, not
    # rounding (measured on B200, side_stream mode: losses equal to 1e-5, worst tensor 1.3e-4)
    for n in s0:
        for k in s0[n]:
            assert rel(s1[n][k], s0[n][k]) < 2e-3, (n, k, rel(s1[n][k], s0[n][k]))


# ------------------------------------------------------------------------------------------------ xVAPitch waveform decoder
def _xvapitch_fixture():
    """Weights, latent and conditioning vector of tests/golden/xvapitch_generator.npz, regenerated from its recorded
    (key, shape) list with the fixture's seeded procedure (tests/golden/make_golden_xvapitch_generator.py)."""
    import ast

    gold = np.load(os.path.join(GOLD, "xvapitch_generator.npz"))
    spec = [(str(k), ast.literal_eval(str(sh))) for k, sh in zip(gold["spec_keys"], gold["spec_shapes"])]
    gen = torch.Generator().manual_seed(7)
    sd = {k: torch.empty(sh) for k, sh in spec}
    for k, sh in spec:
        if k.endswith("weight_v") or k.endswith(".weight"):
            sd[k].copy_(torch.randn(sh, generator=gen) * 0.7 / np.sqrt(sh[1] * sh[2]))
        elif not k.endswith("weight_g"):
            sd[k].copy_((torch.rand(sh, generator=gen) * 2 - 1) * 0.05)
    for k, sh in spec:
        if k.endswith("weight_g"):
            v = sd[k[:-1] + "v"]
            sd[k].copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1) * (1.0 + 0.1 * torch.rand(sh, generator=gen)))
    z = torch.randn(2, 192, 6, generator=gen)
    cond = torch.nn.functional.normalize(torch.randn(2, 512, 1, generator=gen), dim=1)
    assert torch.equal(z, torch.from_numpy(gold["z"])) and torch.equal(cond, torch.from_numpy(gold["g"]))
    return gold, spec, sd, z, cond


def _xvapitch_decoder(lib, sd):
    """The decoder as xvapitch/model.py:134-149 builds it."""
    from xva_trainer_b200 import hifigan as hg

    d = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5], [1, 3, 5], [1, 3, 5]], [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                            inference_padding=0, cond_channels=512, conv_pre_weight_norm=False,
                            conv_post_weight_norm=False, conv_post_bias=False, device="cuda:0")
    res = d.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    d.train()
    return d


def test_xvapitch_decoder_keys_and_forward_match_reference_golden(lib):
    """python/xvapitch/hifigan.py:159-262 as configured at xvapitch/model.py:134-149: same state_dict keys and shapes
    as the reference module's, and its recorded outputs with and without the conditioning vector."""
    gold, spec, sd, z, cond = _xvapitch_fixture()
    d = _xvapitch_decoder(lib, sd)
    assert [(k, tuple(v.shape)) for k, v in d.state_dict().items()] == [(k, tuple(sh)) for k, sh in spec]
    with torch.no_grad():
        y = d(z.cuda(), g=cond.cuda())
        assert y.shape == (2, 1, 1536)
        assert rel(y, torch.from_numpy(gold["y_cond"])) < 2e-3, rel(y, torch.from_numpy(gold["y_cond"]))
        y0 = d(z.cuda())
        assert rel(y0, torch.from_numpy(gold["y_nocond"])) < 2e-3, rel(y0, torch.from_numpy(gold["y_nocond"]))
        assert rel(y, torch.from_numpy(gold["y_nocond"])) > 1e-2       # the conditioning is not a no-op
        assert torch.equal(d.inference(z.cuda()), y0)                  # inference_padding = 0


@pytest.mark.parametrize("with_cond", [True, False])
def test_xvapitch_decoder_backward_matches_oracle(lib, with_cond):
    """Parameter gradients (plain conv_pre / conv_post / cond_layer weights included) and the gradients handed back to
    the latent and the conditioning vector vs autograd through oracle.hifigan.generator_vits; bounds as for the
    HiFi-GAN v1 generator above."""
    gold, spec, sd, z, cond = _xvapitch_fixture()
    d = _xvapitch_decoder(lib, sd)
    gen = torch.Generator().manual_seed(21)
    w = torch.randn(2, 1, 1536, generator=gen)
    y = d(z.cuda(), g=cond.cuda() if with_cond else None)
    d.zero_grad()
    dz, dg = d.backward(w.cuda(), need_input_grad=True)
    torch.cuda.synchronize()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    zl, gl = z.clone().requires_grad_(True), cond.clone().requires_grad_(True)
    yo = ohg.generator_vits(leaves, zl, gl if with_cond else None)
    (yo * w).sum().backward()
    assert rel(y, yo) < 2e-3, rel(y, yo)
    num = den = 0.0
    for k, p in d.named_parameters():
        if k.startswith("cond_layer") and not with_cond:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        e = rel(p.grad, leaves[k].grad)
        assert e < 1e-1, (k, e)
        num += float((p.grad.double().cpu() - leaves[k].grad.double()).pow(2).sum())
        den += float(leaves[k].grad.double().pow(2).sum())
    assert (num / den) ** 0.5 < 4e-2, (num / den) ** 0.5
    # the latent's gradient is the deepest tensor of the backward pass (one more tf32 product than conv_pre's weight
    # gradient): measured 4.7e-2 on this fixture, 1e-4 with the exact-fp32 checker GEMM (wiring test below)
    assert dz.shape == z.shape and rel(dz, zl.grad) < 1e-1, rel(dz, zl.grad)
    if with_cond:
        assert dg.shape == cond.shape and rel(dg, gl.grad) < 1e-1, rel(dg, gl.grad)
    else:
        assert dg is None


def test_xvapitch_decoder_wiring_exact(lib, monkeypatch):
    from xva_trainer_b200 import capi, ops

    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        gold, spec, sd, z, cond = _xvapitch_fixture()
        d = _xvapitch_decoder(lib, sd)
        w = torch.randn(2, 1, 1536, generator=torch.Generator().manual_seed(22))
        y = d(z.cuda(), g=cond.cuda())
        d.zero_grad()
        dz, dg = d.backward(w.cuda(), need_input_grad=True)
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        zl, gl = z.clone().requires_grad_(True), cond.clone().requires_grad_(True)
        yo = ohg.generator_vits(leaves, zl, gl)
        (yo * w).sum().backward()
        assert rel(y, yo) < 1e-5, rel(y, yo)
        # Not 1e-4 as for the v1 fixture above: this fixture has one pre-activation per ~1e5 within 3e-7 of zero in six
        # layers (e.g. the input of resblocks.8.convs1.1), where fp32 summation order decides the sign and with it the
        # leaky-ReLU derivative (1 or 0.1) of that ONE element. scripts/diag_xvapitch_exact.py: worst tensor 1.1e-3
        # (resblocks.8.convs2.0, right behind that element), median 3.5e-4, same against an fp64 oracle; a wiring
        # error is O(1).
        for k, p in d.named_parameters():
            assert rel(p.grad, leaves[k].grad) < 5e-3, (k, rel(p.grad, leaves[k].grad))
        assert rel(dz, zl.grad) < 2e-3 and rel(dg, gl.grad) < 2e-3, (rel(dz, zl.grad), rel(dg, gl.grad))
    finally:
        capi.call("xva_set_operand_rounding", 1)


# ------------------------------------------------------------------------------------------------ xVAPitch discriminator
def _vits_disc(lib):
    from test_oracle_golden import _vits_disc_fixture
    from xva_trainer_b200 import hifigan as hg

    gold, spec, sd, x, x_hat = _vits_disc_fixture()
    m = hg.VitsDiscriminator(device="cuda:0")
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(sh)) for k, sh in spec]
    res = m.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    m.train()
    return hg, m, gold, spec, sd, x, x_hat


def test_vits_discriminator_forward_matches_reference_golden(lib):
    """python/xvapitch/model.py:1590-1631 recorded by tests/golden/make_golden_vits_discriminator.py: state_dict keys,
    the six score maps on the real and the generated waveform (prime length: every period discriminator reflect-pads),
    every feature map against the oracle (itself pinned to the same recording), zero padding channels / rows."""
    hg, m, gold, spec, sd, x, x_hat = _vits_disc(lib)
    with torch.no_grad():
        xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
    ors, ofr, ogs, ofg = ohg.vits_discriminator(sd, x, x_hat)
    B = x.shape[0]
    for i in range(6):
        for got, key in ((xs[i], f"score_real/{i}"), (hs[i], f"score_fake/{i}")):
            want = torch.from_numpy(gold[key])
            got = _to_ref_layout(got, B).reshape(B, -1).cpu()
            assert got.shape == want.shape and rel(got, want) < 3e-3, (key, got.shape, want.shape, rel(got, want))
        for l, (f, of) in enumerate(zip(hf[i], ofg[i])):
            C = of.shape[1]
            full = _to_ref_layout(f, B)
            got = full[:, :C, :of.shape[2]].cpu()
            assert got.shape == of.shape, (i, l, got.shape, of.shape)
            assert rel(got, of) < 3e-3, (i, l, rel(got, of))
            assert float(full[:, C:].abs().max() if full.shape[1] > C else 0.0) == 0.0          # padding channels
            assert float(f[:, of.shape[2]:].abs().max() if f.shape[1] > of.shape[2] else 0.0) == 0.0   # alignment rows


def test_vits_discriminator_step_gradients(lib):
    """D step: LSGAN loss and every parameter gradient (norms recorded from the reference's autograd, full tensors from
    the oracle's); G step: adversarial + feature-matching loss and the gradient wrt the generated waveform, both
    recorded from the reference (xvapitch/losses.py:65-85, 329-342)."""
    hg, m, gold, spec, sd, x, x_hat = _vits_disc(lib)
    B, T = x.shape[0], x.shape[2]
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ors, _, ogs, _ = ohg.vits_discriminator(leaves, x, x_hat)
    ohg.discriminator_loss(ors, ogs).backward()
    m.zero_grad()
    xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
    loss = hg.discriminator_loss_backward(m, xs, hs)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(gold["loss_disc"])) < 2e-3 * float(gold["loss_disc"]), (float(loss), float(gold["loss_disc"]))
    num = den = 0.0
    for (k, p), want_norm in zip(m.named_parameters(), gold["grad_norms"]):
        w = leaves[k].grad
        e = rel(p.grad, w)
        assert e < 6e-2, (k, e)
        assert abs(float(p.grad.double().norm()) - want_norm) < 3e-2 * want_norm, (k, float(p.grad.norm()), want_norm)
        num += float((p.grad.cpu().double() - w.double()).pow(2).sum())
        den += float(w.double().pow(2).sum())
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    for k in ("nets.0.convs.1.weight_v", "nets.0.convs.0.weight_v", "nets.0.convs.0.bias"):       # the 4-channel groups
        # same per-tensor bound as above, now against the reference's own tensors; convs.0 is the deepest (3.4e-2)
        assert rel(dict(m.named_parameters())[k].grad, torch.from_numpy(gold[f"grad/{k}"])) < 6e-2, k
    # ---- G step. The feature-matching gradient back-propagates sign(fake - real) of every feature map, so elements whose
    # difference is within tf32 rounding of zero flip the sign of their seed (about 1 % of them here): measured on B200
    # (scripts/diag_vits_disc_dwave.py) 2.7-4.6 % per net for the adversarial term, up to 19 % for one net's
    # feature-matching term, 2.3-5.1 % on the total depending on the last bit of the packed weights -- and 0.0000 for
    # every net with the exact-fp32 checker GEMM, which is what pins the wiring below.
    from xva_trainer_b200 import capi, ops

    want_gen = torch.from_numpy(gold["dwave_gen"]).reshape(B, T)
    want_all = torch.from_numpy(gold["dwave_gen"] + gold["dwave_feat"]).reshape(B, T)
    xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
    dwave = torch.zeros(B, T, device="cuda")
    lg, lf = hg.generator_adv_loss_backward(m, hs, xf, hf, dwave, pools=0)
    torch.cuda.synchronize()
    assert abs(float(lg) - float(gold["loss_gen"])) < 2e-3 * float(gold["loss_gen"]), (float(lg), float(gold["loss_gen"]))
    assert abs(float(lf) - float(gold["loss_feat"])) < 2e-3 * float(gold["loss_feat"]), (float(lf), float(gold["loss_feat"]))
    assert rel(dwave.cpu(), want_all) < 1.5e-1, rel(dwave.cpu(), want_all)
    xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
    dwave = torch.zeros(B, T, device="cuda")
    hg.generator_adv_loss_backward(m, hs, xf, hf, dwave, pools=0, fm_grad=False)
    assert rel(dwave.cpu(), want_gen) < 6e-2, rel(dwave.cpu(), want_gen)
    orig = ops.gemm_launch
    ops.gemm_launch = lambda args, ref=False: orig(args, True)
    capi.call("xva_set_operand_rounding", 0)
    try:
        xs, xf, hs, hf = m(x.cuda(), x_hat.cuda())
        dwave = torch.zeros(B, T, device="cuda")
        hg.generator_adv_loss_backward(m, hs, xf, hf, dwave, pools=0)
        torch.cuda.synchronize()
        assert rel(dwave.cpu(), want_all) < 1e-3, rel(dwave.cpu(), want_all)
    finally:
        ops.gemm_launch = orig
        capi.call("xva_set_operand_rounding", 1)


def test_spectral_packer_matches_torch(lib):
    """_SnPacker (xva_sn_pack_fwd / _bwd) vs the PyTorch restatement of torch.nn.utils.spectral_norm + re-packing that it
    replaces (_DiscConv.weight / _Disc._packed): two consecutive training calls (the power iteration advances, each call
    keeps its own weights), the u / v buffers after them, eval mode, and the gradient of weight_orig vs autograd."""
    import copy

    from xva_trainer_b200 import hifigan as hg

    msd = hg.MultiScaleDiscriminator(device="cuda:0")
    msd.load_state_dict(ohg.make_disc_state(ohg.msd_spec(), 22))
    d0 = msd.discriminators[0]
    d0.train()
    dref = copy.deepcopy(d0)
    dref.train()
    sn = hg._SnPacker(slots=2)
    d0.register_weights(sn, "0")
    sn.finalize("cuda:0")
    refs = []
    for slot in range(2):
        W = sn.pack(slot, True)
        ref = dref._packed()                       # power iteration on dref's buffers + packed weights under autograd
        refs.append(ref)
        for li, want in enumerate(ref):
            got = W[f"0.{li}"][0]
            assert got.shape == want.shape, (li, got.shape, want.shape)
            assert rel(got, want) < 5e-4, (slot, li, rel(got, want))
        for m, mr in zip(list(d0.convs) + [d0.conv_post], list(dref.convs) + [dref.conv_post]):
            assert rel(m.weight_u, mr.weight_u) < 1e-5 and rel(m.weight_v, mr.weight_v) < 1e-5
    assert not torch.equal(sn.W[0]["0.3"][0], sn.W[1]["0.3"][0])          # the second call used its own sigma
    # backward of the FIRST call (its u, v, sigma, not the buffers' current ones) vs autograd
    gen = torch.Generator(device="cuda").manual_seed(5)
    sn.zero_grads()
    grads = []
    for li, t in enumerate(refs[0]):
        g_ = sn.gW[0][f"0.{li}"][0]
        g_.copy_(torch.randn(g_.shape, device="cuda", generator=gen))
        grads.append(g_.clone())
    params = [m.weight_orig for m in list(dref.convs) + [dref.conv_post]]
    torch.autograd.backward(refs[0], grads)
    for m in list(d0.convs) + [d0.conv_post]:
        m.weight_orig.grad = None
    sn.unpack_grads()
    torch.cuda.synchronize()
    for m, p_ in zip(list(d0.convs) + [d0.conv_post], params):
        assert rel(m.weight_orig.grad, p_.grad) < 1e-4, (tuple(p_.shape), rel(m.weight_orig.grad, p_.grad))
    # eval mode: no power iteration, sigma from the stored vectors
    d0.eval(); dref.eval()
    u_before = d0.convs[2].weight_u.clone()
    W = sn.pack(0, False)
    ref = dref._packed()
    for li, want in enumerate(ref):
        assert rel(W[f"0.{li}"][0], want) < 5e-4, (li, rel(W[f"0.{li}"][0], want))
    assert torch.equal(d0.convs[2].weight_u, u_before)
