"""Host code of the xVAPitch text encoder (xva-trainer_b200/textenc.py) checked WITHOUT a GPU: every C-ABI call it makes
is executed on host memory by tests/cabi_emu.py (the tap-GEMM / softmax / LayerNorm contracts of include/xva_b200.h,
validated in tests/test_cabi_emu_cpu.py; the new element-wise kernels as the very functions of csrc/relattn_body.h compiled
with g++). What is compared: the reference module's recorded outputs (tests/golden/vits_text_encoder.npz), the oracle
(oracle.vits.text_encoder) and its autograd for every parameter gradient -- i.e. which operands, strides, taps, head
slices, padded layouts and saved tensors the module hands to the kernels. What this cannot show -- the CUDA kernels
producing the same numbers on the device -- is tests/test_vits_text_encoder_gpu.py. CPU only."""
import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cabi_emu  # noqa: E402
from oracle import vits as ov  # noqa: E402
from textenc_util import TE_PATCHES, golden_case, oracle_grads, rel, seeded_state  # noqa: E402


def _build(te, sd, layers, lang=12, vocab=50, hidden=192, heads=2, ffn=768, k=3, p=0.0):
    m = te.TextEncoder(vocab, hidden, hidden, ffn, heads, layers, k, p, language_emb_dim=lang, device="cpu")
    m.load_state_dict(sd)
    return m


def test_forward_matches_the_reference_golden_and_state_dict_round_trips():
    g, sd, tokens, lens, lang = golden_case()
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = _build(te, sd, 3)
        back = m.state_dict()
        assert list(back) == list(sd)
        for k_ in sd:
            assert back[k_].shape == sd[k_].shape and torch.equal(back[k_], sd[k_]), k_
        m.eval()
        x, x_emb, mask = m(tokens, lens, lang_emb=lang)
        m_p, logs_p = m(x, lens, stats=True, x_mask=mask)
        used = set(cabi_emu.calls)
    assert {"xva_text_embed_fwd", "xva_rel_band_add", "xva_rel_band_gather", "xva_gemm", "xva_softmax_fwd",
            "xva_layernorm_fwd"} <= used
    assert rel(x_emb, torch.from_numpy(g["x_emb"])) < 1e-6
    assert rel(x, torch.from_numpy(g["x"])) < 2e-5
    assert rel(m_p, torch.from_numpy(g["m_p"])) < 2e-5 and rel(logs_p, torch.from_numpy(g["logs_p"])) < 2e-5
    assert float(x[1, :, 8:].abs().max()) == 0.0 and float(m_p[1, :, 8:].abs().max()) == 0.0
    assert mask.shape == (2, 1, 13) and float(mask.sum()) == 21.0


@pytest.mark.parametrize("T,lens,layers,cfg", [
    (13, [13, 8], 3, dict()),                                   # the golden's shape
    (3, [3, 2], 2, dict()),                                     # shorter than the relative window: the band is clipped
    (37, [37, 20, 33], 2, dict(lang=4, hidden=64, ffn=96, heads=2)),   # 68 channels: another head / pad geometry, T > 32
])
def test_backward_matches_oracle_autograd(T, lens, layers, cfg):
    lang_dim, hidden, vocab = cfg.get("lang", 12), cfg.get("hidden", 192), 50
    sd = seeded_state(layers=layers, lang=lang_dim, hidden=hidden, ffn=cfg.get("ffn", 768), heads=cfg.get("heads", 2))
    gen = torch.Generator().manual_seed(100 + T)
    B = len(lens)
    tokens = torch.randint(1, vocab, (B, T), generator=gen)
    lang = torch.randn(B, lang_dim, 1, generator=gen)
    C = hidden + lang_dim
    rx, rm, rl = torch.randn(B, C, T, generator=gen), torch.randn(B, hidden, T, generator=gen), torch.randn(B, hidden, T, generator=gen)
    re = torch.randn(B, T, hidden, generator=gen) * 0.1
    want_out, want = oracle_grads(sd, tokens, lens, lang, layers, rx, rm, rl, re)

    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = _build(te, sd, layers, lang=lang_dim, hidden=hidden, ffn=cfg.get("ffn", 768), heads=cfg.get("heads", 2))
        m.train()
        m.zero_grad()
        li = torch.tensor(lens, dtype=torch.int32)
        x_cl, x_emb = m.forward_cl(tokens, li, lang.reshape(B, lang_dim).contiguous())
        stats = m.stats_cl(x_cl, li)
        dstats = torch.cat([rm, rl], 1).transpose(1, 2).contiguous()
        dx_stats = m.stats_backward_cl(dstats)
        dlang = m.backward_cl(rx.transpose(1, 2).contiguous() + dx_stats, dx_emb=re)
        got = m.grads()
        used = set(cabi_emu.calls)
    assert {"xva_text_embed_bwd", "xva_pad_cols", "xva_softmax_bwd", "xva_layernorm_bwd", "xva_colsum_items"} <= used
    assert rel(x_cl.transpose(1, 2), want_out["x"]) < 2e-5
    assert rel(stats[..., :hidden].transpose(1, 2), want_out["m_p"]) < 2e-5
    assert rel(dlang.unsqueeze(-1), want["lang"]) < 5e-5
    # (conv_k.bias has no gradient in exact arithmetic -- a key bias shifts every score of a row alike -- so each tensor's
    # error is taken relative to the larger of its own norm and 1e-4 of the largest gradient norm)
    floor = 1e-4 * max(float(want[k_].norm()) for k_ in sd)
    errs = sorted(((float((got[k_] - want[k_]).norm()) / max(float(want[k_].norm()), floor), k_) for k_ in sd), reverse=True)
    assert errs[0][0] < 1e-4, errs[:5]
    # the padded entries of the arena (head padding, relative rows 9..31, pitch columns) received exactly nothing
    V = m._views(m.flat.grad)
    for i in range(layers):
        qw = V[f"l{i}.qkv_w"].view(3, m.num_heads, m.dkp, m.Cp)
        assert float(qw[:, :, m.dk:].abs().max()) == 0.0 and float(qw[..., m.C:].abs().max()) == 0.0
        assert float(V[f"l{i}.ek"][9:].abs().max()) == 0.0 and float(V[f"l{i}.ev"][:, m.dk:].abs().max()) == 0.0
        assert float(V[f"l{i}.o_w"].view(m.C, m.num_heads, m.dkp)[:, :, m.dk:].abs().max()) == 0.0


def _masked_reference(sd, tokens, lens, lang, layers, heads, p, seed, window=4):
    """oracle.vits.text_encoder with the four dropout sites of a layer (glow_tts.py:193, 475, 355, 480) switched on, the
    masks taken from the library's counter hash at the element indices its kernels use: softmax row (h B + b) T + t with
    pitch Tp; GEMM epilogues (b T + t) N + n."""
    import math
    import torch.nn.functional as F
    B, T = tokens.shape
    Ce = sd["emb.weight"].shape[1]
    x_emb = F.embedding(tokens, sd["emb.weight"]) * math.sqrt(Ce)
    x = torch.cat((x_emb, lang.transpose(2, 1).expand(B, T, -1)), dim=-1)             # [B, T, C]
    mask = ov.sequence_mask(lens, T)[:, :, None].to(x.dtype)                           # [B, T, 1]
    x = x * mask
    C = x.shape[2]
    dk, Tp = C // heads, (T + 31) // 32 * 32
    site = [0]

    def scale(shape_idx):
        site[0] += 1
        sd_ = (seed * 0x9E3779B1 + site[0] * 0x85EBCA77) & 0xFFFFFFFFFFFF
        return torch.from_numpy(cabi_emu.dropout_scale(sd_, None, shape_idx, p))

    rows = (np.arange(B)[:, None] * T + np.arange(T)[None, :]).astype(np.uint64)       # b T + t
    epi = lambda N: rows[:, :, None] * np.uint64(N) + np.arange(N, dtype=np.uint64)[None, None, :]
    d = torch.arange(T)[None, :] - torch.arange(T)[:, None]
    near, idx = (d.abs() <= window), (d + window).clamp(0, 2 * window)
    for i in range(layers):
        a = f"encoder.attn_layers.{i}"
        lin = lambda n, t: t @ sd[f"{a}.conv_{n}.weight"][:, :, 0].T + sd[f"{a}.conv_{n}.bias"]
        q, k, v = (lin(n, x).view(B, T, heads, dk).transpose(1, 2) for n in "qkv")      # [B, H, T, dk]
        e_k, e_v = sd[f"{a}.emb_rel_k"][0], sd[f"{a}.emb_rel_v"][0]
        sc = (q @ k.transpose(-2, -1) + torch.einsum("bhid, ijd -> bhij", q, e_k[idx]) * near) / math.sqrt(dk)
        sc = sc.masked_fill((mask.transpose(1, 2) * mask)[:, None] == 0, -1e4)
        pr = F.softmax(sc, dim=-1)
        zrow = (np.arange(heads)[None, :, None] * B + np.arange(B)[:, None, None]) * T + np.arange(T)[None, None, :]   # [B, H, T]
        sidx = zrow.astype(np.uint64)[..., None] * np.uint64(Tp) + np.arange(T, dtype=np.uint64)[None, None, None, :]
        pr = pr * scale(sidx)
        o = pr @ v + torch.einsum("bhij, ijd -> bhid", pr * near, e_v[idx])
        y = lin("o", o.transpose(1, 2).reshape(B, T, C)) * scale(epi(C))
        x = F.layer_norm(x + y, (C,), sd[f"encoder.norm_layers_1.{i}.gamma"], sd[f"encoder.norm_layers_1.{i}.beta"], 1e-5) * mask
        f = f"encoder.ffn_layers.{i}"
        conv = lambda n, t: F.conv1d(t.transpose(1, 2), sd[f"{f}.conv_{n}.weight"], sd[f"{f}.conv_{n}.bias"], padding=1).transpose(1, 2)
        h = torch.relu(conv(1, x))
        h = h * scale(epi(h.shape[2])) * mask
        y = conv(2, h) * scale(epi(C))
        x = F.layer_norm(x + y, (C,), sd[f"encoder.norm_layers_2.{i}.gamma"], sd[f"encoder.norm_layers_2.{i}.beta"], 1e-5) * mask
    return x


def test_dropout_masks_of_forward_and_backward_agree():
    """With dropout on (p = 0.3, four sites per layer) the module's output and every parameter gradient equal those of
    the oracle's arithmetic with the SAME masks applied through autograd -- the masks come from the library's counter
    hash at the indices its kernels use, so a backward that re-derived a different mask than the forward drew, or a site
    with the wrong seed, shows up as an O(1) gradient error."""
    layers, T, lens, heads = 2, 9, [9, 6], 2
    sd = seeded_state(layers=layers, lang=4, hidden=32, ffn=64, heads=heads)
    gen = torch.Generator().manual_seed(5)
    tokens = torch.randint(1, 50, (2, T), generator=gen)
    lang = torch.randn(2, 4, 1, generator=gen)
    r = torch.randn(2, T, 36, generator=gen)
    li = torch.tensor(lens, dtype=torch.int32)
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = te.TextEncoder(50, 32, 32, 64, heads, layers, 3, 0.3, language_emb_dim=4, device="cpu")
        m.load_state_dict(sd)
        m.train()
        m.zero_grad()
        x, _ = m.forward_cl(tokens, li, lang.reshape(2, 4).contiguous())
        x_again, _ = m.forward_cl(tokens, li, lang.reshape(2, 4).contiguous())
        assert torch.equal(x, x_again)                          # same counter, same masks
        dlang = m.backward_cl(r.clone())
        got = m.grads()
        m.eval()
        x_eval, _ = m.forward_cl(tokens, li, lang.reshape(2, 4).contiguous())
        m.train()
        m.step_dropout()
        x_next, _ = m.forward_cl(tokens, li, lang.reshape(2, 4).contiguous())
        seed = m.seed
    assert rel(x, x_eval) > 0.05 and not torch.equal(x, x_next)   # dropout was on; the counter draws new masks
    p = {k_: v.clone().requires_grad_(True) for k_, v in sd.items()}
    lg = lang.clone().requires_grad_(True)
    want_x = _masked_reference(p, tokens, lens, lg, layers, heads, 0.3, seed)
    (want_x * r).sum().backward()
    assert rel(x, want_x) < 2e-5
    assert rel(dlang.unsqueeze(-1), lg.grad) < 1e-4
    floor = 1e-4 * max(float(v.grad.norm()) for k_, v in p.items() if v.grad is not None)
    for k_, v in p.items():
        if v.grad is None:                                      # proj.*: not on this path
            continue
        assert float((got[k_] - v.grad).norm()) / max(float(v.grad.norm()), floor) < 2e-4, k_


# ------------------------------------------------------------------------------------------------ pitch predictor
def _pitch_case():
    from textenc_util import fill_pitch, pitch_ref_spec

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vits_pitch_predictor.npz"))
    gen = torch.Generator().manual_seed(71)
    sd = fill_pitch(pitch_ref_spec(), gen)
    x = torch.randn(2, 13, 196, generator=gen)
    spk = torch.nn.functional.normalize(torch.randn(2, 512, 1, generator=gen), dim=1)
    r = torch.randn(2, 1, 13, generator=gen)
    assert torch.equal(x, torch.from_numpy(g["x"])) and torch.equal(r, torch.from_numpy(g["r"]))
    return g, sd, x, spk, r


def test_pitch_predictor_matches_the_reference_golden_and_oracle_autograd():
    """textenc.RelativePositioningPitchEnergyEncoder (xvapitch/model.py:1268-1356 as built at :154-168: 196 + 512 = 708
    channels, 3 layers, out_channels = 1) through the emulated C ABI: output vs the reference recording, every gradient vs the
    oracle's autograd and the reference's recorded gradient norms; the six parameters the reference never trains are state,
    not optimizer parameters; state_dict round trip in the reference's keys."""
    g, sd, x, spk, r = _pitch_case()
    lens = [13, 8]
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ov.pitch_predictor(p, x, lens, spk, num_layers=3)
    (want * r).sum().backward()
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = te.RelativePositioningPitchEnergyEncoder(1, 196, 768, 2, 3, 3, 0.0, conditioning_emb_dim=512, device="cpu")
        m.load_state_dict(sd)
        back = m.state_dict()
        assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
        n_dead = sum(int(np.prod(sd[k].shape)) for k in m.dead_keys())
        assert len(m.dead_keys()) == 6 and m.flat.numel() < sum(v.numel() for v in sd.values()) * 1.2 and n_dead > 1_000_000
        m.train()
        m.zero_grad()
        pred = m(x, lens, speaker_emb=spk)
        dxs = m.backward(r, need_input_grad=True)
        got = m.grads()
        used = set(cabi_emu.calls)
    assert {"xva_gemm", "xva_layernorm_fwd", "xva_layernorm_bwd", "xva_rel_band_add", "xva_pad_cols", "xva_colsum_items"} <= used
    assert pred.shape == (2, 1, 13)
    assert rel(pred, torch.from_numpy(g["pitch_pred"])) < 2e-5 and rel(pred, want.detach()) < 2e-5
    assert float(pred[1, :, 8:].abs().max()) == 0.0
    assert set(got) == set(str(k) for k in g["has_grad"])            # exactly the parameters the reference gives a gradient
    floor = 1e-4 * max(float(p[k].grad.norm()) for k in got)
    for k in got:
        assert float((got[k] - p[k].grad).norm()) / max(float(p[k].grad.norm()), floor) < 1e-4, k
        assert abs(float(got[k].norm()) - float(g["gnorm/" + k])) <= 3e-4 * max(float(g["gnorm/" + k]), 10 * floor), k
    # input gradients (not needed in training: the reference detaches x) against autograd of the oracle
    xr, sr = x.clone().requires_grad_(True), spk.clone().requires_grad_(True)
    (ov.pitch_predictor(sd, xr, lens, sr, num_layers=3) * r).sum().backward()
    assert rel(dxs[0], xr.grad) < 1e-4 and rel(dxs[1].unsqueeze(-1), sr.grad) < 1e-4


def test_pitch_predictor_big_model_shape():
    """The other configuration xVAPitch builds (big = 1: 256 + 12 + 512 = 780 channels, heads of 390 padded to 416)."""
    from textenc_util import fill_pitch, pitch_ref_spec

    gen = torch.Generator().manual_seed(3)
    sd = fill_pitch(pitch_ref_spec(layers=2, hidden=268), gen)
    x = torch.randn(2, 9, 268, generator=gen)
    spk = torch.randn(2, 512, 1, generator=gen)
    r = torch.randn(2, 1, 9, generator=gen)
    lens = [9, 4]
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = ov.pitch_predictor(p, x, lens, spk, num_layers=2)
    (want * r).sum().backward()
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = te.RelativePositioningPitchEnergyEncoder(1, 268, 768, 2, 2, 3, 0.0, conditioning_emb_dim=512, device="cpu")
        m.load_state_dict(sd)
        m.train()
        m.zero_grad()
        pred = m(x, lens, speaker_emb=spk)
        m.backward(r)
        got = m.grads()
    assert (m.C, m.Cp, m.dk, m.dkp) == (780, 800, 390, 416)
    assert rel(pred, want.detach()) < 2e-5
    floor = 1e-4 * max(float(p[k].grad.norm()) for k in got)
    for k in got:
        assert float((got[k] - p[k].grad).norm()) / max(float(p[k].grad.norm()), floor) < 1e-4, k


def test_adamw_over_the_flat_arenas_matches_torch_and_leaves_pad_and_dead_entries_alone():
    """hifigan.AdamW([module.flat]) -- the optimizer of the xVAPitch generator (training_util.py:56-57: lr 1.75e-4, betas
    0.8 / 0.99, eps 1e-9, weight decay 0.01) -- on the text encoder and the pitch predictor: after one step every
    reference-shaped parameter equals torch.optim.AdamW's on the oracle's gradients, pad entries of the arena are still
    exactly zero, and the pitch predictor's six untrained tensors are bit-identical (torch skips parameters without a
    gradient, weight decay included)."""
    from textenc_util import fill_pitch, pitch_ref_spec

    gen = torch.Generator().manual_seed(71)
    sd = fill_pitch(pitch_ref_spec(layers=2), gen)
    x = torch.randn(2, 11, 196, generator=gen)
    spk = torch.randn(2, 512, 1, generator=gen)
    r = torch.randn(2, 1, 11, generator=gen)
    lens = [11, 7]
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    topt = torch.optim.AdamW(list(p.values()), lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
    (ov.pitch_predictor(p, x, lens, spk, num_layers=2) * r).sum().backward()
    topt.step()
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        hg = cabi_emu.load_module("hifigan", [('if dev.type != "cuda":', "if False:")])
        m = te.RelativePositioningPitchEnergyEncoder(1, 196, 768, 2, 2, 3, 0.0, conditioning_emb_dim=512, device="cpu")
        m.load_state_dict(sd)
        opt = hg.AdamW([m.flat], lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
        opt.zero_grad()
        m.train()
        m(x, lens, speaker_emb=spk)
        m.backward(r)
        opt.step()
        after = m.state_dict()
        V = m._views(m.flat.data)
        for i in range(2):
            qw = V[f"l{i}.qkv_w"].view(3, m.num_heads, m.dkp, m.Cp)
            assert float(qw[:, :, m.dk:].abs().max()) == 0.0 and float(qw[..., m.C:].abs().max()) == 0.0
            assert float(V[f"l{i}.ek"][9:].abs().max()) == 0.0 and float(V[f"l{i}.ev"][:, m.dk:].abs().max()) == 0.0
        assert float(V["proj_w"][:, m.C:].abs().max()) == 0.0
    for k in sd:
        if k in m.dead_keys():
            assert torch.equal(after[k], sd[k]) and p[k].grad is None, k
        else:
            # Adam's first step is lr * g / |g|: where a gradient is zero in exact arithmetic (the key bias, and the key
            # weights' 512 speaker-embedding columns -- constant over the keys of an utterance, so softmax-invariant) both
            # sides take a full-size step in the direction of their own rounding noise. Compare where the gradient is real.
            gr = p[k].grad
            real = gr.abs() > 1e-3 * gr.abs().max()
            if k.endswith("conv_k.bias"):
                continue
            assert float(real.float().mean()) > 0.2, k
            assert rel((after[k] - sd[k])[real], (p[k].detach() - sd[k])[real]) < 2e-3, k


def test_tf32_operand_model_reproduces_the_error_measured_on_the_b200():
    """cabi_emu.TF32 = True models the product path's operand precision (xva_gemm truncates fp32 operands to tf32, producers
    round to nearest). On the text encoder's 13-token case the error against the fp32 oracle it predicts on the CPU is the
    one MEASURED on the B200 (profiles/r02_textenc.json: x 5.94e-4, gradient vector 6.88e-3, worst tensor 1.82e-2) to within
    a few per cent -- which is what lets the bounds of GPU tests that have not run yet be set from a prediction
    (profiles/r02_tf32_parity_predicted.txt)."""
    import json
    import math

    measured = json.load(open(os.path.join(ROOT, "profiles", "r02_textenc.json")))["parity_product_path"][0]
    assert measured["case"].startswith("T=13")
    layers, lens = 3, [13, 8]
    sd = seeded_state(layers=layers)
    gen = torch.Generator().manual_seed(100 + 13)
    tokens = torch.randint(1, 50, (2, 13), generator=gen)
    lang = torch.randn(2, 12, 1, generator=gen)
    rx, rm, rl = torch.randn(2, 204, 13, generator=gen), torch.randn(2, 192, 13, generator=gen), torch.randn(2, 192, 13, generator=gen)
    re = torch.randn(2, 13, 192, generator=gen) * 0.1
    want_out, want = oracle_grads(sd, tokens, lens, lang, layers, rx, rm, rl, re)
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        m = _build(te, sd, layers)
        m.train()
        m.zero_grad()
        cabi_emu.TF32 = True
        try:
            li = torch.tensor(lens, dtype=torch.int32)
            x_cl, _ = m.forward_cl(tokens, li, lang.reshape(2, 12).contiguous())
            m.stats_cl(x_cl, li)
            dx = m.stats_backward_cl(torch.cat([rm, rl], 1).transpose(1, 2).contiguous())
            m.backward_cl(rx.transpose(1, 2).contiguous() + dx, dx_emb=re)
        finally:
            cabi_emu.TF32 = False
        got = m.grads()
    ex = rel(x_cl.transpose(1, 2), want_out["x"])
    num = sum(float((got[k] - want[k]).norm()) ** 2 for k in sd)
    den = sum(float(want[k].norm()) ** 2 for k in sd)
    eg = math.sqrt(num / den)
    assert abs(ex - measured["x"]) < 0.05 * measured["x"], (ex, measured["x"])
    assert abs(eg - measured["grad_global"]) < 0.05 * measured["grad_global"], (eg, measured["grad_global"])


def test_pitch_predictor_gpu_checks_hold_under_the_tf32_operand_model(monkeypatch):
    """The pitch predictor's GPU tests have not run on hardware yet (tests/test_vits_zz_pitch_predictor_gpu.py). Their probe
    (tests/pitch_predictor_gpu_probe.py) is executed here on the CPU instead, through the emulated C ABI with the tf32 operand
    model on -- the model that reproduces the text encoder's measured errors -- and has to meet the very bounds the GPU tests
    assert: exact-checker wiring 2e-5 / 2e-4, product path 3e-3 / 2e-2 / 6e-2, reference golden 3e-3, AMP-free AdamW step
    leaving the six untrained tensors bit-identical."""
    import json

    import pitch_predictor_gpu_probe as P
    import xva_trainer_b200

    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    lines = []
    monkeypatch.setattr("builtins.print", lambda *a, **k: lines.append(" ".join(str(x) for x in a)))
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        hg = cabi_emu.load_module("hifigan", [('if dev.type != "cuda":', "if False:")])
        for name, mod in (("textenc", te), ("hifigan", hg)):
            monkeypatch.setitem(sys.modules, f"xva_trainer_b200.{name}", mod)
            monkeypatch.setattr(xva_trainer_b200, name, mod, raising=False)
        cabi_emu.TF32 = True
        try:
            P.main()
        finally:
            cabi_emu.TF32 = False
            cabi_emu.ROUNDING_ON = True
    out = json.loads([ln for ln in lines if ln.startswith("PITCH_PREDICTOR_PROBE ")][-1].split(" ", 1)[1])
    for e in out["exact"]:
        assert e["same_keys"] and e["fwd"] < 2e-5 and e["grad_worst"] < 2e-4, e
    for e in out["product"]:
        assert e["same_keys"] and e["pad_max"] == 0.0 and 1e-4 < e["fwd"] < 3e-3 and e["grad_global"] < 2e-2 and e["grad_worst"] < 6e-2, e
    assert 1e-4 < out["golden_fwd"] < 3e-3
    assert out["keys_in_reference_order"] and out["dead_untouched"] and out["moved"] == out["trainable"]
