"""Generates tests/golden/xvapitch_generator.npz: the UNMODIFIED xVAPitch waveform decoder (python/xvapitch/hifigan.py
HifiganGenerator, configured as xvapitch/model.py:134-149) on seeded inputs, with its seeded-and-perturbed state dict.
Groundwork for SURVEY.md section 8f rank 1. Build container only:   python tests/golden/make_golden_xvapitch_generator.py
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
np.bool = bool                                 # xvapitch/util.py:28 uses the alias numpy >= 1.24 removed
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()
from python.xvapitch.hifigan import HifiganGenerator  # noqa: E402


def fill_state(params, spec, gen):
    """Seeded weights at magnitudes where every layer matters (the default init is N(0, 0.01)-like and tiny). Two passes in
    the module's parameter order: everything but the weight-norm gains, then the gains from the new directions. The test
    repeats this procedure from the recorded (key, shape) list instead of storing 14.6 M parameters."""
    with torch.no_grad():
        for k, shape in spec:
            if k.endswith("weight_v") or k.endswith(".weight"):
                params[k].copy_(torch.randn(shape, generator=gen) * 0.7 / np.sqrt(shape[1] * shape[2]))
            elif not k.endswith("weight_g"):
                params[k].copy_((torch.rand(shape, generator=gen) * 2 - 1) * 0.05)
        for k, shape in spec:
            if k.endswith("weight_g"):
                v = params[k[:-1] + "v"]
                params[k].copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1) * (1.0 + 0.1 * torch.rand(shape, generator=gen)))


def main():
    torch.manual_seed(1234)
    G = HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2], inference_padding=0,
                         cond_channels=512, conv_pre_weight_norm=False, conv_post_weight_norm=False, conv_post_bias=False)
    gen = torch.Generator().manual_seed(7)
    named = list(G.named_parameters())
    fill_state({k: p for k, p in named}, [(k, tuple(p.shape)) for k, p in named], gen)
    G.eval()
    z = torch.randn(2, 192, 6, generator=gen)
    g = torch.nn.functional.normalize(torch.randn(2, 512, 1, generator=gen), dim=1)
    out = {"spec_keys": np.array([k for k, _ in named]), "spec_shapes": np.array([str(tuple(p.shape)) for _, p in named])}
    out["z"], out["g"] = z.numpy(), g.numpy()
    with torch.no_grad():
        out["y_cond"] = G(z, g=g).numpy()
        out["y_nocond"] = G(z).numpy()
    np.savez_compressed(os.path.join(HERE, "xvapitch_generator.npz"), **out)
    print({k: v.shape for k, v in out.items()}, len(named), "parameter tensors")


if __name__ == "__main__":
    main()
