"""Shared by the CPU (emulated C ABI) and GPU parity tests of the xVAPitch text encoder: the seeded state of
tests/golden/make_golden_vits_text_encoder.py, the golden case, and the oracle's outputs / autograd gradients."""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

TE_PATCHES = [('if dev.type != "cuda":', "if False:"),
              ('dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
               'dev = torch.device("cpu")')]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ref_spec(layers, lang=12, hidden=192, ffn=768, heads=2, vocab=50, k=3, window=4):
    """(key, shape) of TextEncoder(...).named_parameters() in the reference's registration order."""
    C = hidden + lang
    dk = C // heads
    spec = [("emb.weight", (vocab, hidden))]
    for i in range(layers):
        a = f"encoder.attn_layers.{i}"
        spec += [(f"{a}.emb_rel_k", (1, 2 * window + 1, dk)), (f"{a}.emb_rel_v", (1, 2 * window + 1, dk))]
        for n in "qkvo":
            spec += [(f"{a}.conv_{n}.weight", (C, C, 1)), (f"{a}.conv_{n}.bias", (C,))]
    for i in range(layers):
        spec += [(f"encoder.norm_layers_1.{i}.gamma", (C,)), (f"encoder.norm_layers_1.{i}.beta", (C,))]
    for i in range(layers):
        f = f"encoder.ffn_layers.{i}"
        spec += [(f"{f}.conv_1.weight", (ffn, C, k)), (f"{f}.conv_1.bias", (ffn,)), (f"{f}.conv_2.weight", (C, ffn, k)),
                 (f"{f}.conv_2.bias", (C,))]
    for i in range(layers):
        spec += [(f"encoder.norm_layers_2.{i}.gamma", (C,)), (f"encoder.norm_layers_2.{i}.beta", (C,))]
    spec += [("proj.weight", (2 * hidden, C, 1)), ("proj.bias", (2 * hidden,))]
    return spec


def fill(spec, gen, hidden=192):
    """The seeded fill of make_golden_vits_text_encoder.py (non-trivial LayerNorm parameters and biases)."""
    sd = {}
    for k, sh in spec:
        if k.endswith("gamma"):
            sd[k] = 1.0 + 0.1 * torch.randn(sh, generator=gen)
        elif k.endswith("beta") or k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(sh, generator=gen)
        elif k == "emb.weight":
            sd[k] = torch.randn(sh, generator=gen) * hidden ** -0.5
        elif "emb_rel" in k:
            sd[k] = torch.randn(sh, generator=gen) * 96 ** -0.5
        else:
            sd[k] = torch.randn(sh, generator=gen) * 0.7 / np.sqrt(int(np.prod(sh[1:])))
    return sd


def seeded_state(layers, lang=12, hidden=192, ffn=768, heads=2, seed=61):
    return fill(ref_spec(layers, lang, hidden, ffn, heads), torch.Generator().manual_seed(seed), hidden)


def golden_case():
    """-> (npz, state dict, tokens [2, 13], lens [13, 8], lang [2, 12, 1]) exactly as the golden script drew them."""
    g = np.load(os.path.join(HERE, "golden", "vits_text_encoder.npz"))
    spec = [(str(k), ast.literal_eval(str(sh))) for k, sh in zip(g["spec_keys"], g["spec_shapes"])]
    assert spec == ref_spec(3), "the reference's parameter order / shapes differ from textenc_util.ref_spec"
    gen = torch.Generator().manual_seed(61)
    sd = fill(spec, gen)
    tokens = torch.randint(1, 50, (2, 13), generator=gen)
    lens = [13, 8]
    lang = torch.randn(2, 12, 1, generator=gen)
    assert torch.equal(tokens, torch.from_numpy(g["tokens"])) and torch.equal(lang, torch.from_numpy(g["lang_emb"]))
    return g, sd, tokens, lens, lang


def oracle_grads(sd, tokens, lens, lang, layers, rx, rm, rl, re):
    """Oracle forward + autograd of  L = <x, rx> + <m_p, rm> + <logs_p, rl> + <x_emb, re>  (x [B, C, T], m_p / logs_p
    [B, out, T], x_emb [B, T, hidden]) -> ({x, x_emb, m_p, logs_p}, {key: gradient} + {"lang": dL/d lang_emb [B, L, 1]})."""
    from oracle import vits as ov

    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lg = lang.clone().requires_grad_(True)
    x, x_emb, mask = ov.text_encoder(p, tokens, lens, lg, num_layers=layers)
    m_p, logs_p = ov.text_encoder_stats(p, x, mask)
    loss = (x * rx).sum() + (m_p * rm).sum() + (logs_p * rl).sum() + (x_emb * re).sum()
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    grads["lang"] = lg.grad
    return {"x": x.detach(), "x_emb": x_emb.detach(), "m_p": m_p.detach(), "logs_p": logs_p.detach()}, grads


# ------------------------------------------------------------------------------------------------ pitch predictor
def fill_pitch(spec, gen):
    """Seeded fill of the pitch predictor's parameters (make_golden_vits_pitch_predictor.py): non-trivial LayerNorm
    parameters and biases, weights scaled by 0.7 / sqrt(fan_in)."""
    sd = {}
    for k, sh in spec:
        if k.endswith("gamma"):
            sd[k] = 1.0 + 0.1 * torch.randn(sh, generator=gen)
        elif k.endswith("beta") or k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(sh, generator=gen)
        elif "emb_rel" in k:
            sd[k] = torch.randn(sh, generator=gen) * 96 ** -0.5
        else:
            sd[k] = torch.randn(sh, generator=gen) * 0.7 / np.sqrt(int(np.prod(sh[1:])))
    return sd


def pitch_ref_spec(layers=3, hidden=196, cond=512, ffn=768, heads=2, k=3, window=4):
    """(key, shape) of RelativePositioningPitchEnergyEncoder(out_channels=1, ...).named_parameters() in the reference's
    registration order: the last layer's FFN / LayerNorm are built with out_channels = 1, encoder.proj after attn_layers[-1]."""
    C = hidden + cond
    dk = C // heads
    spec = []
    for i in range(layers):
        a = f"encoder.attn_layers.{i}"
        spec += [(f"{a}.emb_rel_k", (1, 2 * window + 1, dk)), (f"{a}.emb_rel_v", (1, 2 * window + 1, dk))]
        for n in "qkvo":
            spec += [(f"{a}.conv_{n}.weight", (C, C, 1)), (f"{a}.conv_{n}.bias", (C,))]
    for i in range(layers):
        spec += [(f"encoder.norm_layers_1.{i}.gamma", (C,)), (f"encoder.norm_layers_1.{i}.beta", (C,))]
    for i in range(layers):
        f, o = f"encoder.ffn_layers.{i}", (C if i + 1 < layers else 1)
        spec += [(f"{f}.conv_1.weight", (ffn, C, k)), (f"{f}.conv_1.bias", (ffn,)), (f"{f}.conv_2.weight", (o, ffn, k)),
                 (f"{f}.conv_2.bias", (o,))]
    for i in range(layers):
        o = C if i + 1 < layers else 1
        spec += [(f"encoder.norm_layers_2.{i}.gamma", (o,)), (f"encoder.norm_layers_2.{i}.beta", (o,))]
    spec += [("encoder.proj.weight", (1, C, 1)), ("encoder.proj.bias", (1,))]
    return spec


def fill_sdp(spec, gen):
    """Seeded fill of the stochastic duration predictor's parameters (make_golden_vits_sdp.py): the reference zero-initialises
    every ConvFlow projection (identity splines); here they get small random values so the splines are exercised."""
    sd = {}
    for k, sh in spec:
        if k.endswith("gamma"):
            sd[k] = 1.0 + 0.1 * torch.randn(sh, generator=gen)
        elif k.endswith("beta") or k.endswith("bias") or k.endswith("translation") or k.endswith("log_scale"):
            sd[k] = 0.1 * torch.randn(sh, generator=gen)
        else:
            sd[k] = torch.randn(sh, generator=gen) * 0.7 / np.sqrt(int(np.prod(sh[1:])))
    return sd
