"""Generates tests/golden/vits_alignment.npz: the alignment / prior-expansion / KL part of xVAPitch.train_step, run from the
UNMODIFIED reference:

  * python/xvapitch/model.py:763-776 (log-likelihood of every latent frame under every text token's prior, then
    maximum_path) and :855-856 (prior expansion) are statements inside train_step, not callable on their own: this
    script reads those source lines from the reference file at generation time, dedents them and executes them on seeded
    tensors with the reference's own maximum_path (python/xvapitch/util.py:14-53). Nothing of them is stored here.
  * VitsGeneratorLoss.kl_loss (python/xvapitch/losses.py:86-103) is called directly, with autograd for its gradients.

SURVEY.md section 8f rank 1. Build container only:   python tests/golden/make_golden_vits_alignment.py"""
import math
import os
import sys
import textwrap
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install_xvapitch()
from python.xvapitch.losses import VitsGeneratorLoss  # noqa: E402
from python.xvapitch.util import maximum_path  # noqa: E402

MODEL = os.path.join(_ref_import.REFERENCE_ROOT, "python", "xvapitch", "model.py")


def reference_lines(first, last, must_contain):
    lines = open(MODEL).read().splitlines()[first - 1:last]
    src = textwrap.dedent("\n".join(lines))
    for s in must_contain:
        assert s in src, f"model.py:{first}-{last} no longer holds {s!r}: re-derive the line numbers"
    return src


def main():
    gen = torch.Generator().manual_seed(53)
    B, C, Tt, Ts = 3, 64, 23, 61
    x_lens, y_lens = torch.tensor([23, 17, 9]), torch.tensor([61, 48, 30])
    x_mask = (torch.arange(Tt)[None, :] < x_lens[:, None]).float().unsqueeze(1)        # [B, 1, Tt]
    y_mask = (torch.arange(Ts)[None, :] < y_lens[:, None]).float().unsqueeze(1)        # [B, 1, Ts]
    m_p = torch.randn(B, C, Tt, generator=gen) * x_mask
    logs_p = 0.3 * torch.randn(B, C, Tt, generator=gen) * x_mask
    z_p = (torch.randn(B, C, Ts, generator=gen) * 1.2) * y_mask
    logs_q = 0.3 * torch.randn(B, C, Ts, generator=gen) * y_mask
    ns = {"torch": torch, "math": math, "maximum_path": maximum_path, "x_mask_d": x_mask, "y_mask": y_mask.clone(),
          "m_p": m_p.clone(), "logs_p": logs_p.clone(), "z_p": z_p.clone()}
    # keep logp alive for the recording: drop the reference's `del` statements, nothing else
    src = reference_lines(763, 776, ["attn_mask = torch.unsqueeze(x_mask_d, -1)", "o_scale = torch.exp(-2 * logs_p)",
                                     "attn = maximum_path(logp, attn_mask.squeeze(1))"])
    src = "\n".join(ln for ln in src.splitlines() if not ln.strip().startswith("del "))
    exec(src, ns)
    logp, attn = ns["logp"], ns["attn"]
    ns2 = {"torch": torch, "attn": attn, "m_p": m_p.clone(), "logs_p": logs_p.clone()}
    exec(reference_lines(855, 856, ['m_p = torch.einsum("klmn, kjm -> kjn", [attn, m_p])',
                                    'logs_p = torch.einsum("klmn, kjm -> kjn", [attn, logs_p])']), ns2)
    m_pe, logs_pe = ns2["m_p"], ns2["logs_p"]
    leaves = [t.clone().requires_grad_(True) for t in (z_p, logs_q, m_pe, logs_pe)]
    loss_kl, _ = VitsGeneratorLoss.kl_loss(*leaves, y_mask)
    loss_kl.backward()
    out = {"x_lens": x_lens.numpy(), "y_lens": y_lens.numpy(), "m_p": m_p.numpy(), "logs_p": logs_p.numpy(), "z_p": z_p.numpy(),
           "logs_q": logs_q.numpy(), "logp": logp.numpy(), "attn": attn.squeeze(1).numpy().astype(np.int8),
           "durations": attn.sum(3).squeeze(1).numpy(), "m_p_expanded": m_pe.numpy(), "logs_p_expanded": logs_pe.numpy(),
           "loss_kl": np.float64(loss_kl.item())}
    for name, t in zip(("z_p", "logs_q", "m_p_expanded", "logs_p_expanded"), leaves):
        out[f"grad/{name}"] = t.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "vits_alignment.npz"), **out)
    print("loss_kl", loss_kl.item(), "durations", attn.sum(3).squeeze(1)[1].tolist())


if __name__ == "__main__":
    main()
