"""xVAPitch --hifi_only path (xva-trainer_b200/vits.py + hifigan.HifiganGenerator / VitsDiscriminator, all math through
libxva_b200.so) vs the CPU oracle (oracle/vits.py, pinned to the reference by tests/test_oracle_golden.py) and vs the
step recorded from the unmodified reference modules (tests/golden/vits_hifi_only.npz).

Tolerances as for the HiFi-GAN path (tests/test_hifigan_gpu.py): tf32 tensor-core operands rounded to nearest, fp32
accumulation; element-wise kernels 1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import hifigan as ohg
from oracle import vits as ov

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_gated_activation_and_sample_kernels(lib):
    from xva_trainer_b200 import capi, ops

    capi.call("xva_set_operand_rounding", 0)
    try:
        g = torch.Generator().manual_seed(3)
        B, T, H = 3, 37, 64
        x_in = torch.randn(B, T, 2 * H, generator=g).requires_grad_(True)
        d = torch.randn(B, T, H, generator=g)
        want = torch.tanh(x_in[..., :H]) * torch.sigmoid(x_in[..., H:])
        (want * d).sum().backward()
        got = ops.gated_act(x_in.detach().cuda(), H)
        assert rel(got, want) < 1e-6
        assert rel(ops.gated_act_bwd(d.cuda(), x_in.detach().cuda(), H), x_in.grad) < 1e-6
        # posterior sample
        C = 32
        stats = (torch.randn(B, T, 2 * C, generator=g) * 0.5).requires_grad_(True)
        eps = torch.randn(B, T, C, generator=g)
        lens = torch.tensor([37, 20, 1], dtype=torch.int32)
        mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(-1)
        z = (stats[..., :C] + eps * torch.exp(stats[..., C:])) * mask
        dz = torch.randn(B, T, C, generator=g)
        (z * dz).sum().backward()
        got = ops.vits_sample(stats.detach().cuda(), eps.cuda(), lens.cuda())
        assert rel(got, z) < 1e-6 and float(got[1, 20:].abs().max()) == 0.0
        assert rel(ops.vits_sample_bwd(dz.cuda(), eps.cuda(), stats.detach().cuda(), lens.cuda()), stats.grad) < 1e-6
        # per-item column sums, strided input and output views
        x = torch.randn(B, T, 96, generator=g)
        out = torch.ones(B, 200)
        want = out.clone()
        want[:, 40:104] += x[..., 16:80].sum(1)
        xo, oo = x.cuda(), out.cuda()
        ops.colsum_items_(xo[..., 16:80], oo[:, 40:104])
        assert rel(oo, want) < 1e-6
    finally:
        capi.call("xva_set_operand_rounding", 1)


def test_vits_mel_matches_reference_and_autograd(lib):
    """MelSpectrogram.vits() = TorchSTFT(1024, 256, 1024, ..., use_mel, do_amp_to_db) of xvapitch/audio.py:138-181: the
    log-mels recorded from the reference's own criterion.stft, and its backward vs autograd through the oracle."""
    from xva_trainer_b200 import hifigan as hg

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vits_hifi_only.npz"))
    mel = hg.MelSpectrogram.vits(device="cuda:0")
    for wkey, mkey in (("wav_seg", "mel_real"), ("o", "mel_fake")):
        w = torch.from_numpy(gold[wkey]).reshape(2, -1)
        got = mel(w.cuda()).transpose(1, 2).cpu()
        want = torch.from_numpy(gold[mkey])
        assert got.shape == want.shape and rel(got, want) < 2e-3, (mkey, got.shape, rel(got, want))
    w = torch.from_numpy(gold["o"]).reshape(2, -1).clone().requires_grad_(True)
    d = torch.randn(2, 80, 33, generator=torch.Generator().manual_seed(5))
    (ov.torch_stft_mel(w) * d).sum().backward()
    mel(w.detach().cuda())
    got = mel.backward(d.transpose(1, 2).contiguous().cuda())
    assert rel(got, w.grad) < 5e-3, rel(got, w.grad)


@pytest.mark.parametrize("with_g", [True, False])
def test_wn_matches_oracle(lib, with_g):
    """WN (python/xvapitch/wavenet.py:16-106) stand-alone: output, input / conditioning gradients and every parameter
    gradient vs autograd through oracle.vits.wn, ragged mask."""
    from xva_trainer_b200 import vits

    H, L, K, C = 64, 3, 5, 64
    m = vits.WN(H, H, K, 1, L, c_in_channels=C).to("cuda:0")
    gen = torch.Generator().manual_seed(9)
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
    B, T = 2, 45
    x = torch.randn(B, H, T, generator=gen)
    g = torch.nn.functional.normalize(torch.randn(B, C, 1, generator=gen), dim=1) if with_g else None
    lens = [45, 31]
    mask = ov.sequence_mask(lens, T)[:, None, :].float()
    d = torch.randn(B, H, T, generator=gen)
    leaves = {f"enc.{k}": v.clone().requires_grad_(True) for k, v in sd.items()}
    xl = x.clone().requires_grad_(True)
    gl = g.clone().requires_grad_(True) if with_g else None
    want = ov.wn(leaves, "enc", xl, mask, gl, num_layers=L, hidden=H, kernel=K)
    (want * d).sum().backward()
    got = m(x.cuda(), mask.cuda(), g.cuda() if with_g else None)
    assert rel(got, want) < 2e-3, rel(got, want)
    m.zero_grad()
    dx, dg = m.backward(d.cuda())
    torch.cuda.synchronize()
    assert rel(dx, xl.grad) < 5e-3, rel(dx, xl.grad)
    if with_g:
        assert rel(dg, gl.grad) < 5e-3, rel(dg, gl.grad)
    for k, p in m.named_parameters():
        if k.startswith("cond_layer") and not with_g:
            continue
        assert rel(p.grad, leaves[f"enc.{k}"].grad) < 1e-2, (k, rel(p.grad, leaves[f"enc.{k}"].grad))


def _modules(lib):
    from test_oracle_golden import _vits_hifi_only_fixture
    from xva_trainer_b200 import hifigan as hg
    from xva_trainer_b200 import vits

    gold, specs, sds, linear, waveform, d_vectors = _vits_hifi_only_fixture()
    enc = vits.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device="cuda:0")
    dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2], inference_padding=0,
                              cond_channels=512, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                              conv_post_bias=False, device="cuda:0")
    disc = hg.VitsDiscriminator(device="cuda:0")
    for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
        assert [(k, tuple(v.shape)) for k, v in mod.state_dict().items()] == [(k, tuple(sh)) for k, sh in specs[name]], name
        res = mod.load_state_dict(sds[name])
        assert not res.missing_keys and not res.unexpected_keys
        mod.train()
    return gold, specs, sds, linear, waveform, d_vectors, enc, dec, disc


def test_posterior_encoder_matches_reference_golden_and_oracle(lib):
    """python/xvapitch/model.py:1422-1475 (513 -> 192, 16 WaveNet layers, conditioning 512): z / mean / log_scale recorded
    from the reference module with its own N(0, 1) draw replayed, and every parameter gradient of sum(z * w) vs autograd
    through the oracle."""
    gold, specs, sds, linear, waveform, d_vectors, enc, dec, disc = _modules(lib)
    g = torch.nn.functional.normalize(d_vectors).unsqueeze(-1)
    eps = torch.from_numpy(gold["eps"])
    lens = [int(v) for v in gold["y_lengths"]]
    z, m_q, logs_q, mask = enc(linear.cuda(), lens, g=g.cuda(), eps=eps.cuda())
    for got, key in ((z, "z"), (m_q, "m_q"), (logs_q, "logs_q")):
        want = torch.from_numpy(gold[key])
        assert got.shape == want.shape and rel(got, want) < 3e-3, (key, got.shape, rel(got, want))
    assert float(z[1, :, 35:].abs().max()) == 0.0 and mask.shape == (2, 1, 40) and float(mask.sum()) == 75.0
    w = torch.randn(2, 192, 40, generator=torch.Generator().manual_seed(4))
    leaves = {k: v.clone().requires_grad_(True) for k, v in sds["enc"].items()}
    zo, _, _, _ = ov.posterior_encoder(leaves, linear, lens, g, eps)
    (zo * w).sum().backward()
    enc.zero_grad()
    enc.backward(w.cuda())
    torch.cuda.synchronize()
    num = den = 0.0
    for k, p in enc.named_parameters():
        e = rel(p.grad, leaves[k].grad)
        assert e < 3e-2, (k, e)
        num += float((p.grad.double().cpu() - leaves[k].grad.double()).pow(2).sum())
        den += float(leaves[k].grad.double().pow(2).sum())
    assert (num / den) ** 0.5 < 1e-2, (num / den) ** 0.5


def test_hifi_only_step_matches_reference_golden_and_oracle(lib):
    """One --hifi_only iteration (vits.HifiOnlyStep) with the recorded random draws replayed: every loss of the step
    recorded from the reference modules within 2e-3, the segment starts exactly, and all 447 updated parameter tensors
    within 5e-3 of the oracle's (bound as in test_full_hifigan_step_matches_oracle: the first AdamW step is sign-like)
    with the recorded norms of the updated tensors and of their change."""
    from xva_trainer_b200 import vits

    gold, specs, sds, linear, waveform, d_vectors, enc, dec, disc = _modules(lib)
    before = {n: {k: v.clone() for k, v in sd.items()} for n, sd in sds.items()}
    step = vits.HifiOnlyStep(enc, dec, disc)
    eps, u = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["u"])
    lens = [int(v) for v in gold["y_lengths"]]
    losses = step.step(linear, lens, waveform, d_vectors, eps=eps, u=u)
    torch.cuda.synchronize()
    assert losses["slice_ids"].tolist() == gold["slice_ids"].tolist()
    for k in ("loss", "loss_gen", "loss_feat", "loss_mel", "loss_disc"):
        a, b = float(losses[k]), float(gold[k])
        assert abs(a - b) < 2e-3 * abs(b), (k, a, b)
    want = ov.hifi_only_step(sds["enc"], sds["dec"], sds["disc"], linear, waveform, d_vectors, lens, eps, u, {})
    for k in ("loss", "loss_gen", "loss_feat", "loss_mel", "loss_disc"):
        assert abs(float(losses[k]) - want[k]) < 2e-3 * abs(want[k]), (k, float(losses[k]), want[k])
    for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
        after = mod.state_dict()
        moved = 0
        for i, (k, _) in enumerate(specs[name]):
            assert rel(after[k], sds[name][k]) < 5e-3, (name, k, rel(after[k], sds[name][k]))
            moved += int(not torch.equal(after[k].cpu(), before[name][k]))
            n_after = float(after[k].double().norm())
            assert abs(n_after - gold[f"{name}/after_norms"][i]) < 1e-3 * gold[f"{name}/after_norms"][i] + 1e-9, (name, k)
            delta = float((after[k].double().cpu() - before[name][k].double()).norm())
            assert abs(delta - gold[f"{name}/delta_norms"][i]) < 0.1 * gold[f"{name}/delta_norms"][i] + 1e-9, (name, k, delta)
        assert moved == len(specs[name]), (name, moved)


def test_flow_matches_reference_golden_and_oracle(lib):
    """vits.ResidualCouplingBlocks = python/xvapitch/model.py:1358-1421 (4 flows x 4-layer WN): forward and reverse outputs
    recorded from the reference module, reverse(forward(x)) = x, input / conditioning gradients as recorded from its
    autograd, every parameter gradient vs the oracle's and its recorded norm."""
    from test_oracle_golden import _vits_flow_fixture
    from xva_trainer_b200 import vits

    gold, spec, sd = _vits_flow_fixture()
    flow = vits.ResidualCouplingBlocks(192, 192, kernel_size=5, dilation_rate=1, num_layers=4, cond_channels=512, device="cuda:0")
    assert [(k, tuple(v.shape)) for k, v in flow.state_dict().items()] == [(k, tuple(sh)) for k, sh in spec]
    res = flow.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    flow.train()
    x, cond, w = (torch.from_numpy(gold[k]) for k in ("x", "g", "w"))
    lens = [int(v) for v in gold["lens"]]
    mask = ov.sequence_mask(lens, x.shape[2])[:, None, :].float()
    z_p = flow(x.cuda(), mask.cuda(), g=cond.cuda())
    assert rel(z_p, torch.from_numpy(gold["z_p"])) < 2e-3, rel(z_p, torch.from_numpy(gold["z_p"]))
    flow.zero_grad()
    dx, dg = flow.backward(w.cuda())
    torch.cuda.synchronize()
    assert rel(dx, torch.from_numpy(gold["dx"])) < 5e-3, rel(dx, torch.from_numpy(gold["dx"]))
    assert rel(dg, torch.from_numpy(gold["dg"])) < 5e-3, rel(dg, torch.from_numpy(gold["dg"]))
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (ov.residual_coupling_blocks(leaves, x, mask, cond) * w).sum().backward()
    for (k, p), want_norm in zip(flow.named_parameters(), gold["grad_norms"]):
        assert rel(p.grad, leaves[k].grad) < 2e-2, (k, rel(p.grad, leaves[k].grad))
        assert abs(float(p.grad.double().norm()) - want_norm) < 1e-2 * want_norm + 1e-12, (k, float(p.grad.norm()), want_norm)
    with torch.no_grad():
        back = flow(z_p, mask.cuda(), g=cond.cuda(), reverse=True)
    assert rel(back, torch.from_numpy(gold["reverse_of_z_p"])) < 2e-3
    assert rel(back, x) < 2e-3, rel(back, x)


def test_hifi_only_step_at_bench_shape(lib):
    """The workload bench.py measures (batch 16 x 256 spectrogram frames, ragged, 8192-sample segments), one iteration
    from the same seeded state on both sides with the same random draws: every loss within 2e-3 of the oracle's, all 447
    updated tensors within 5e-3, and the size-independent properties of the step -- masked latent rows are exactly zero,
    every segment start is inside its utterance, all parameters move."""
    import bench
    from xva_trainer_b200 import hifigan as hg, vits

    B, T = 16, 256
    gen = torch.Generator().manual_seed(77)
    mk = lambda spec, scale: {k: v for k, v in _seeded_state(spec, gen, scale).items()}
    sds = {"enc": mk(ov.posterior_encoder_spec(), 0.7), "dec": mk(ov.decoder_spec(), 0.7), "disc": mk(ohg.vits_disc_spec(), 1.0)}
    enc = vits.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device="cuda:0")
    dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2], inference_padding=0,
                              cond_channels=512, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                              conv_post_bias=False, device="cuda:0")
    disc = hg.VitsDiscriminator(device="cuda:0")
    for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
        res = mod.load_state_dict(sds[name])
        assert not res.missing_keys and not res.unexpected_keys
        mod.train()
    linear, lens, waveform, d_vectors = bench.synthetic_vits_batch(B, T, 1)
    eps, u = torch.randn(B, 192, T, generator=gen), torch.rand(B, generator=gen)
    step = vits.HifiOnlyStep(enc, dec, disc)
    before = {n: {k: v.clone() for k, v in sd.items()} for n, sd in sds.items()}
    losses = step.step(linear, lens.tolist(), waveform, d_vectors, eps=eps, u=u)
    torch.cuda.synchronize()
    starts = losses["slice_ids"].tolist()
    assert all(0 <= s and s + 32 <= int(n) for s, n in zip(starts, lens.tolist()))
    z = enc.z_cl
    for b, n in enumerate(lens.tolist()):
        assert float(z[b, n:].abs().max() if n < T else 0.0) == 0.0
    want = ov.hifi_only_step(sds["enc"], sds["dec"], sds["disc"], linear, waveform, d_vectors, lens.tolist(), eps, u, {})
    for k in ("loss", "loss_gen", "loss_feat", "loss_mel", "loss_disc"):
        assert abs(float(losses[k]) - want[k]) < 2e-3 * abs(want[k]), (k, float(losses[k]), want[k])
    for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
        after = mod.state_dict()
        for k, v in sds[name].items():
            assert rel(after[k], v) < 5e-3, (name, k, rel(after[k], v))
            assert not torch.equal(after[k].cpu(), before[name][k]), (name, k)


def _seeded_state(spec, gen, scale):
    sd = {k: torch.empty(sh) for k, sh in spec}
    for k, sh in spec:
        if k.endswith("weight_v") or k.endswith(".weight"):
            sd[k].copy_(torch.randn(sh, generator=gen) * scale / np.sqrt(int(np.prod(sh[1:]))))
        elif not k.endswith("weight_g"):
            sd[k].copy_((torch.rand(sh, generator=gen) * 2 - 1) * 0.05)
    for k, sh in spec:
        if k.endswith("weight_g"):
            v = sd[k[:-1] + "v"]
            sd[k].copy_(v.flatten(1).norm(dim=1).view(sh) * (1.0 + 0.1 * torch.rand(sh, generator=gen)))
    return sd


def test_prior_alignment_and_kl_match_reference_golden(lib):
    """vits.prior_alignment / kl_loss vs the recording made from the reference's own source lines
    (tests/golden/make_golden_vits_alignment.py: model.py:763-777, 855-856, losses.py:86-103): log-likelihoods to fp32
    rounding, the path and the durations BIT-EXACT, the expanded prior exact, the KL loss and its four gradients; and
    the expansion's backward vs autograd through the oracle's einsum."""
    from xva_trainer_b200 import vits

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vits_alignment.npz"))
    t = lambda k: torch.from_numpy(g[k])
    r = vits.prior_alignment(t("z_p").cuda(), t("m_p").cuda(), t("logs_p").cuda(), g["x_lens"], g["y_lens"])
    torch.cuda.synchronize()
    xl, yl = g["x_lens"], g["y_lens"]
    for b in range(len(xl)):       # compared where the search reads them (the reference multiplies by the mask first)
        got = r["logp"][b, :yl[b], :xl[b]].cpu().transpose(0, 1)
        want = t("logp")[b, :xl[b], :yl[b]]
        assert float((got - want).abs().max()) < 2e-4 * float(want.abs().max()), b
    assert np.array_equal(r["attn"].squeeze(1).cpu().numpy().astype(np.int8), g["attn"])
    assert np.array_equal(r["durations"].squeeze(1).cpu().numpy(), g["durations"])
    assert torch.equal(r["m_p"].cpu(), t("m_p_expanded")) and torch.equal(r["logs_p"].cpu(), t("logs_p_expanded"))
    loss, grads = vits.kl_loss(t("z_p").cuda(), t("logs_q").cuda(), r["m_p"], r["logs_p"], yl)
    assert abs(float(loss) - float(g["loss_kl"])) < 1e-5 * abs(float(g["loss_kl"]))
    for got, k in zip(grads, ("z_p", "logs_q", "m_p_expanded", "logs_p_expanded")):
        assert rel(got, t(f"grad/{k}")) < 1e-5, (k, rel(got, t(f"grad/{k}")))
    # backward of the expansion
    gen = torch.Generator().manual_seed(3)
    dm, dl = torch.randn(t("m_p_expanded").shape, generator=gen), torch.randn(t("m_p_expanded").shape, generator=gen)
    ml, ll = t("m_p").clone().requires_grad_(True), t("logs_p").clone().requires_grad_(True)
    path = t("attn").float()
    (torch.einsum("bts, bct -> bcs", path, ml) * dm).sum().backward()
    (torch.einsum("bts, bct -> bcs", path, ll) * dl).sum().backward()
    got_m, got_l = vits.prior_expand_backward(dm.cuda(), dl.cuda(), r["cum"], t("m_p").shape[2])
    assert rel(got_m, ml.grad) < 1e-6 and rel(got_l, ll.grad) < 1e-6


def test_hifi_only_step_same_with_streams(lib):
    """The stream-level parallelism the step uses inside a captured graph (hifigan._Side: one weight-gradient stream per
    launching stream; hifigan._Branches: the six discriminator nets / the three ResBlocks of an MRF stage side by side),
    forced on for eagerly launched steps: the same two iterations from the same state give the same losses and weights up
    to the order of the fp32 atomic additions of the split weight gradients (bounds as in
    tests/test_hifigan_gpu.py::test_two_stream_backward_gives_the_same_step)."""
    from xva_trainer_b200 import hifigan as hg, vits

    gold, specs, sds, linear, waveform, d_vectors, *_ = _modules(lib)
    eps, u = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["u"])
    lens = [int(v) for v in gold["y_lengths"]]
    results = []
    was = (hg._Side.enabled, hg._Branches.n, hg._Branches.n_gen)
    try:
        for flag in (False, True):
            hg._Side.enabled = flag
            hg._Branches.n = 8 if flag else 0
            hg._Branches.n_gen = 3 if flag else 0
            _, _, _, _, _, _, enc, dec, disc = _modules(lib)
            step = vits.HifiOnlyStep(enc, dec, disc)
            for _ in range(2):
                losses = step.step(linear, lens, waveform, d_vectors, eps=eps, u=u)
            torch.cuda.synchronize()
            results.append(({k: float(v) for k, v in losses.items() if k != "slice_ids"},
                            {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("enc", enc), ("dec", dec), ("disc", disc))}))
    finally:
        hg._Side.enabled, hg._Branches.n, hg._Branches.n_gen = was
    (l0, s0), (l1, s1) = results
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 1e-4 * abs(l0[k]) + 1e-7, (k, l0[k], l1[k])
    for n in s0:
        for k in s0[n]:
            assert rel(s1[n][k], s0[n][k]) < 2e-3, (n, k, rel(s1[n][k], s0[n][k]))
