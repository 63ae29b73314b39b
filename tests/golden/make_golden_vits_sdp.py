"""Generates tests/golden/vits_sdp.npz: the UNMODIFIED xVAPitch stochastic duration predictor
(python/xvapitch/sdp.py:179-300 StochasticDurationPredictor over DilatedDepthSeparableConv / ElementwiseAffine / ConvFlow
and the rational-quadratic splines of python/xvapitch/util.py:206-399; built at model.py:123-131) in its training direction
on seeded inputs: the negative log-likelihood per utterance and the autograd gradient norm of every parameter, with the
module's own N(0, 1) draw recorded (same seed, same first call). Oracle groundwork for SURVEY.md section 8f rank 1 -- the
engine-side module is not built. Build container only:
    python tests/golden/make_golden_vits_sdp.py"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _ref_import  # noqa: E402

_ref_import.install_xvapitch()
from python.xvapitch.sdp import StochasticDurationPredictor  # noqa: E402
from textenc_util import fill_sdp  # noqa: E402


def main():
    torch.manual_seed(1234)
    C, Lang, cond = 192, 4, 512
    sdp = StochasticDurationPredictor(C, C, 3, 0.5, 4, cond_channels=cond, language_emb_dim=Lang)        # model.py:123-131
    named = list(sdp.named_parameters())
    spec = [(k, tuple(p.shape)) for k, p in named]
    gen = torch.Generator().manual_seed(81)
    sd = fill_sdp(spec, gen)
    with torch.no_grad():
        for k, p in named:
            p.copy_(sd[k])
    sdp.eval()                                               # dropout (p = 0.5 inside the two condition encoders) off
    B, T = 2, 11
    lens = torch.tensor([11, 7])
    mask = (torch.arange(T)[None, :] < lens[:, None]).float()[:, None, :]
    x = torch.randn(B, C + Lang, T, generator=gen) * mask
    dr = torch.randint(1, 9, (B, 1, T), generator=gen).float() * mask
    g = torch.nn.functional.normalize(torch.randn(B, cond, 1, generator=gen), dim=1)
    lang = torch.randn(B, Lang, 1, generator=gen)
    torch.manual_seed(77)
    noise = torch.randn(B, 2, T)
    torch.manual_seed(77)
    nll = sdp(x, mask, dr=dr, g=g, lang_emb=lang)
    (nll / mask.sum([1, 2])).sum().backward()                # the duration loss of losses.py: nll / sum(mask), summed
    out = {"spec_keys": np.array([k for k, _ in spec]), "spec_shapes": np.array([str(sh) for _, sh in spec]),
           "lens": lens.numpy(), "x": x.numpy(), "dr": dr.numpy(), "g": g.numpy(), "lang": lang.numpy(), "noise": noise.numpy(),
           "nll": nll.detach().numpy(), "grad_norms": np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in named])}
    np.savez_compressed(os.path.join(HERE, "vits_sdp.npz"), **out)
    print(len(named), "tensors", sum(p.numel() for _, p in named), "parameters; nll", nll.detach().numpy(),
          "no grad:", [k for k, p in named if p.grad is None])


if __name__ == "__main__":
    main()
