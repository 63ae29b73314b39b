"""xVAPitch text encoder (xva-trainer_b200/textenc.py + csrc/relattn.cu, all math through libxva_b200.so) on the device vs
the recording of the unmodified reference module (tests/golden/vits_text_encoder.npz) and vs the CPU oracle
(oracle.vits.text_encoder, pinned to that recording by tests/test_oracle_golden.py) with autograd for the gradients.

Three layers of evidence, as for the FastPitch step (tests/test_fastpitch_gpu.py):
  (1) the five new element-wise kernels one by one against torch, exact;
  (2) WIRING, exact arithmetic: every tap-GEMM routed to the fp32 checker kernel with operand rounding off -- forward
      <= 2e-5, every parameter gradient <= 2e-4 of the oracle's;
  (3) the PRODUCT path (tcgen05 tap-GEMM, tf32 operands rounded to nearest). Measured on B200 over the four cases below
      (profiles/r02_textenc.txt, scripts/prof_textenc.py): forward 4.9e-4 .. 6.2e-4, d(lang_emb) 7.7e-3 .. 1.2e-2, gradient
      vector 6.9e-3 .. 9.4e-3, worst tensor 1.8e-2 .. 2.9e-2 (an FFN conv_1 weight / bias; relative to the larger of its
      own norm and 1 % of the largest gradient norm), median tensor 4e-3 .. 7.6e-3 -- the level of the FastPitch FFT
      blocks at their toy shapes (9e-4 / 9e-3 / 3.4e-2, profiles/r02_parity_table.txt). Bounds = 2 x the measured maxima.
The host code of the module is additionally checked on the CPU against the same oracle (tests/test_vits_text_encoder_cpu.py)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from textenc_util import golden_case, oracle_grads, rel, seeded_state  # noqa: E402

pytestmark = pytest.mark.gpu


def test_element_kernels(lib):
    from xva_trainer_b200 import capi, ops

    capi.call("xva_set_operand_rounding", 0)
    try:
        g = torch.Generator().manual_seed(3)
        B, T, C, L, V, ld = 3, 21, 20, 4, 11, 32
        tokens = torch.randint(0, V, (B, T), generator=g)
        emb, lang = torch.randn(V, C, generator=g), torch.randn(B, L, generator=g)
        lens = torch.tensor([21, 9, 1], dtype=torch.int32)
        mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(-1)
        x, x_emb = ops.text_embed(tokens.cuda(), emb.cuda(), lang.cuda(), lens.cuda(), 1.5, ld)
        want = torch.zeros(B, T, ld)
        want[..., :C] = emb[tokens] * 1.5
        want[..., C:C + L] = lang[:, None, :]
        assert torch.equal(x.cpu(), want * mask) and torch.equal(x_emb.cpu(), emb[tokens] * 1.5)
        # embedding backward: a scatter-add over repeated tokens, live rows only
        dout = torch.randn(B, T, ld, generator=g)
        demb = torch.zeros(V, C).cuda()
        ops.text_embed_bwd_(tokens.cuda(), dout.cuda(), lens.cuda(), C, 1.5, demb)
        wd = torch.zeros(V, C).index_add_(0, tokens.reshape(-1), (dout[..., :C] * mask * 1.5).reshape(-1, C))
        assert rel(demb, wd) < 1e-6
        ops.text_embed_bwd_(tokens.cuda(), dout.cuda(), None, C, 1.0, demb)            # lens = NULL: every row
        wd.index_add_(0, tokens.reshape(-1), dout[..., :C].reshape(-1, C))
        assert rel(demb, wd) < 1e-6
        # band add / gather, window 4, T = 21 and T = 3 (shorter than the window), padded pitches
        for T_ in (21, 3):
            Z, W, ldp = 4, 4, 32
            s = torch.randn(Z, T_, ldp, generator=g)
            r = torch.randn(Z, T_, 32, generator=g)
            d = torch.arange(T_)[None, :] - torch.arange(T_)[:, None]
            near, idx = (d.abs() <= W), (d + W).clamp(0, 2 * W)
            want = s.clone()
            want[..., :T_] += torch.gather(r, 2, idx[None].expand(Z, T_, T_)) * near
            got = ops.rel_band_add_(s.clone().cuda(), r.cuda(), T_, W)
            assert torch.equal(got.cpu(), want)
            gat = ops.rel_band_gather(s.cuda(), T_, W, 32).cpu()
            wg = torch.zeros(Z, T_, 32)
            for t in range(T_):
                for rr in range(2 * W + 1):
                    j = t + rr - W
                    if 0 <= j < T_:
                        wg[:, t, rr] = s[:, t, j]
            assert torch.equal(gat, wg)
        src = torch.randn(5, 7, 204, generator=g)
        pad = ops.pad_cols(src.cuda(), 224).cpu()
        assert torch.equal(pad[..., :204], src) and float(pad[..., 204:].abs().max()) == 0.0
    finally:
        capi.call("xva_set_operand_rounding", 1)


def _build(sd, layers, lang=12, hidden=192, heads=2, ffn=768, p=0.0):
    from xva_trainer_b200 import textenc

    m = textenc.TextEncoder(50, hidden, hidden, ffn, heads, layers, 3, p, language_emb_dim=lang)
    m.load_state_dict(sd)
    return m


def _run(m, tokens, lens, lang, rx, rm, rl, re):
    dev = m.flat.device
    B, hidden = tokens.shape[0], m.hidden_channels
    m.train()
    m.zero_grad()
    li = torch.tensor(lens, dtype=torch.int32, device=dev)
    x_cl, x_emb = m.forward_cl(tokens.to(dev), li, lang.reshape(B, -1).contiguous().to(dev))
    stats = m.stats_cl(x_cl, li)
    dx_stats = m.stats_backward_cl(torch.cat([rm, rl], 1).transpose(1, 2).contiguous().to(dev))
    dlang = m.backward_cl(rx.transpose(1, 2).contiguous().to(dev) + dx_stats, dx_emb=re.to(dev))
    torch.cuda.synchronize()
    out = {"x": x_cl.transpose(1, 2), "x_emb": x_emb, "m_p": stats[..., :hidden].transpose(1, 2),
           "logs_p": stats[..., hidden:].transpose(1, 2)}
    return out, m.grads(), dlang.unsqueeze(-1)


def _case(T, lens, layers, cfg, seed):
    lang_dim, hidden = cfg.get("lang", 12), cfg.get("hidden", 192)
    sd = seeded_state(layers=layers, lang=lang_dim, hidden=hidden, ffn=cfg.get("ffn", 768), heads=cfg.get("heads", 2))
    gen = torch.Generator().manual_seed(seed)
    B = len(lens)
    tokens = torch.randint(1, 50, (B, T), generator=gen)
    lang = torch.randn(B, lang_dim, 1, generator=gen)
    C = hidden + lang_dim
    rx, rm, rl = torch.randn(B, C, T, generator=gen), torch.randn(B, hidden, T, generator=gen), torch.randn(B, hidden, T, generator=gen)
    re = torch.randn(B, T, hidden, generator=gen) * 0.1
    return sd, tokens, lang, (rx, rm, rl, re)


CASES = [(13, [13, 8], 3, dict()), (3, [3, 2], 2, dict()), (70, [70, 41, 64, 5], 2, dict()),
         (37, [37, 20, 33], 2, dict(lang=4, hidden=64, ffn=96, heads=2))]


@pytest.mark.parametrize("T,lens,layers,cfg", CASES)
def test_wiring_exact_with_fp32_checker_gemm(lib, monkeypatch, T, lens, layers, cfg):
    from xva_trainer_b200 import capi, ops

    sd, tokens, lang, seeds = _case(T, lens, layers, cfg, 100 + T)
    want_out, want = oracle_grads(sd, tokens, lens, lang, layers, *seeds)
    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        m = _build(sd, layers, **{k: v for k, v in cfg.items()})
        out, got, dlang = _run(m, tokens, lens, lang, *seeds)
    finally:
        capi.call("xva_set_operand_rounding", 1)
    for k in ("x", "x_emb", "m_p", "logs_p"):
        assert rel(out[k], want_out[k]) < 2e-5, k
    assert rel(dlang, want["lang"]) < 1e-4
    floor = 1e-4 * max(float(want[k].norm()) for k in sd)
    for k in sd:
        assert float((got[k].cpu() - want[k]).norm()) / max(float(want[k].norm()), floor) < 2e-4, k


@pytest.mark.parametrize("T,lens,layers,cfg", CASES)
def test_product_path_matches_the_oracle(lib, T, lens, layers, cfg):
    sd, tokens, lang, seeds = _case(T, lens, layers, cfg, 100 + T)
    want_out, want = oracle_grads(sd, tokens, lens, lang, layers, *seeds)
    m = _build(sd, layers, **{k: v for k, v in cfg.items()})
    out, got, dlang = _run(m, tokens, lens, lang, *seeds)
    assert rel(out["x_emb"], want_out["x_emb"]) < 1e-6
    for k in ("x", "m_p", "logs_p"):
        assert rel(out[k], want_out[k]) < 1.3e-3, k
    pad = out["x"].cpu()
    for b, n in enumerate(lens):
        assert float(pad[b, :, n:].abs().max()) == 0.0 if n < T else True
    assert rel(dlang, want["lang"]) < 2.4e-2
    floor = 1e-2 * max(float(want[k].norm()) for k in sd)
    num = sum(float((got[k].cpu() - want[k]).norm()) ** 2 for k in sd)
    den = sum(float(want[k].norm()) ** 2 for k in sd)
    assert math.sqrt(num / den) < 2e-2
    for k in sd:
        assert float((got[k].cpu() - want[k]).norm()) / max(float(want[k].norm()), floor) < 6e-2, k
    # pad entries of the arena never receive a gradient (AdamW leaves them zero)
    V = m._views(m.flat.grad)
    for i in range(layers):
        qw = V[f"l{i}.qkv_w"].view(3, m.num_heads, m.dkp, m.Cp)
        assert float(qw[:, :, m.dk:].abs().max()) == 0.0 and float(qw[..., m.C:].abs().max()) == 0.0
        assert float(V[f"l{i}.ek"][9:].abs().max()) == 0.0 and float(V[f"l{i}.ev"][:, m.dk:].abs().max()) == 0.0


def test_forward_matches_the_reference_golden(lib):
    g, sd, tokens, lens, lang = golden_case()
    m = _build(sd, 3)
    m.eval()
    x, x_emb, mask = m(tokens, lens, lang_emb=lang)
    m_p, logs_p = m(x, lens, stats=True, x_mask=mask)
    assert rel(x_emb, torch.from_numpy(g["x_emb"])) < 1e-6
    assert rel(x, torch.from_numpy(g["x"])) < 1.3e-3
    assert rel(m_p, torch.from_numpy(g["m_p"])) < 1.3e-3 and rel(logs_p, torch.from_numpy(g["logs_p"])) < 1.3e-3
    assert float(x[1, :, 8:].abs().max()) == 0.0
    back = m.state_dict()
    assert list(back) == list(sd) and all(torch.equal(back[k].cpu(), sd[k]) for k in sd)


def test_dropout_is_reproducible_and_advances_with_the_counter(lib):
    sd, tokens, lang, seeds = _case(13, [13, 8], 2, dict(), 7)
    m = _build(sd, 2, p=0.1)
    m.train()
    dev = m.flat.device
    li = torch.tensor([13, 8], dtype=torch.int32, device=dev)
    args = (tokens.to(dev), li, lang.reshape(2, -1).contiguous().to(dev))
    x1, _ = m.forward_cl(*args)
    x2, _ = m.forward_cl(*args)
    assert torch.equal(x1, x2)
    m.step_dropout()
    x3, _ = m.forward_cl(*args)
    assert not torch.equal(x1, x3) and rel(x3, x1) < 1.0
    m.eval()
    x4, _ = m.forward_cl(*args)
    assert rel(x4, x1) > 1e-3


def test_adamw_keeps_the_pad_entries_zero_and_moves_the_rest(lib):
    """One optimizer step of the flat arena with hifigan.AdamW (the optimizer of the xVAPitch generator,
    python/xvapitch/training_util.py:56-57): real entries move, pad entries (zero value, zero gradient) stay exactly zero."""
    from xva_trainer_b200 import hifigan

    sd, tokens, lang, seeds = _case(13, [13, 8], 2, dict(), 9)
    m = _build(sd, 2)
    opt = hifigan.AdamW([m.flat], lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    opt.zero_grad()
    dev = m.flat.device
    m.train()
    li = torch.tensor([13, 8], dtype=torch.int32, device=dev)
    x_cl, _ = m.forward_cl(tokens.to(dev), li, lang.reshape(2, -1).contiguous().to(dev))
    m.backward_cl(seeds[0].transpose(1, 2).contiguous().to(dev))
    opt.step()
    torch.cuda.synchronize()
    after = m.state_dict()
    moved = [k for k in before if not torch.equal(before[k], after[k])]
    assert len(moved) >= len(before) - 6                       # everything but proj.* (not on this path) and conv_k.bias-like zeros
    V = m._views(m.flat.data)
    for i in range(2):
        qw = V[f"l{i}.qkv_w"].view(3, m.num_heads, m.dkp, m.Cp)
        assert float(qw[:, :, m.dk:].abs().max()) == 0.0 and float(qw[..., m.C:].abs().max()) == 0.0
        assert float(V[f"l{i}.ek"][9:].abs().max()) == 0.0
