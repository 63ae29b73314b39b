"""Generates tests/golden/vits_pitch_predictor.npz: the UNMODIFIED xVAPitch pitch predictor
(python/xvapitch/model.py:1268-1356 RelativePositioningPitchEnergyEncoder, built at model.py:154-168: the text encoder's
output concatenated with the 512-channel speaker embedding, 3 layers of RelativePositionTransformer with out_channels = 1 --
so the last layer's FFN and LayerNorm are built but their result is discarded, glow_tts.py:476-483, and a 1x1 projection to
one channel takes their place) on seeded inputs, dropout off, with the autograd gradients of sum(pitch_pred * r) for every
parameter that receives one (norm + the first entries). SURVEY.md section 8f rank 1. Build container only:
    python tests/golden/make_golden_vits_pitch_predictor.py"""
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
from python.xvapitch.model import RelativePositioningPitchEnergyEncoder  # noqa: E402
from textenc_util import fill_pitch  # noqa: E402


def main():
    torch.manual_seed(1234)
    hidden, cond, layers = 196, 512, 3                       # model.py:154-168 with the small model's 192 + 4 channels
    enc = RelativePositioningPitchEnergyEncoder(out_channels=1, hidden_channels=hidden, hidden_channels_ffn=768, num_heads=2,
                                                num_layers=layers, kernel_size=3, dropout_p=0.1, conditioning_emb_dim=cond)
    named = list(enc.named_parameters())
    spec = [(k, tuple(p.shape)) for k, p in named]
    gen = torch.Generator().manual_seed(71)
    sd = fill_pitch(spec, gen)
    with torch.no_grad():
        for k, p in named:
            p.copy_(sd[k])
    enc.eval()
    B, T = 2, 13
    lens = torch.tensor([13, 8])
    x = torch.randn(B, T, hidden, generator=gen)
    spk = torch.nn.functional.normalize(torch.randn(B, cond, 1, generator=gen), dim=1)
    r = torch.randn(B, 1, T, generator=gen)
    pred = enc(x, lens, speaker_emb=spk)
    (pred * r).sum().backward()
    out = {"spec_keys": np.array([k for k, _ in spec]), "spec_shapes": np.array([str(sh) for _, sh in spec]),
           "lens": lens.numpy(), "x": x.numpy(), "spk": spk.numpy(), "r": r.numpy(), "pitch_pred": pred.detach().numpy()}
    has_grad = []
    for k, p in named:
        if p.grad is not None:
            has_grad.append(k)
            out["gnorm/" + k] = np.array(float(p.grad.norm()))
            out["ghead/" + k] = p.grad.reshape(-1)[:8].numpy().copy()
    out["has_grad"] = np.array(has_grad)
    np.savez_compressed(os.path.join(HERE, "vits_pitch_predictor.npz"), **out)
    print(len(named), "tensors,", len(has_grad), "with gradients,", sum(p.numel() for _, p in named), "parameters; pred", tuple(pred.shape))
    print("without gradient:", [k for k, _ in named if k not in has_grad])


if __name__ == "__main__":
    main()
