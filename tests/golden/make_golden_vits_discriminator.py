"""Generates tests/golden/vits_discriminator.npz: the UNMODIFIED xVAPitch discriminator (python/xvapitch/model.py:1548-1631
VitsDiscriminator = one DiscriminatorS with VITS channel widths + DiscriminatorP for periods 2, 3, 5, 7, 11 from
xvapitch/hifigan.py:301-370) on seeded waveforms, with the reference's own loss functions (xvapitch/losses.py:65-85,
329-342) and autograd. SURVEY.md section 8f rank 1. Build container only:
    python tests/golden/make_golden_vits_discriminator.py
The 46.7 M parameters are not stored: the test regenerates them from the recorded (key, shape) list with fill_state()."""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install_xvapitch()
from python.xvapitch.losses import VitsDiscriminatorLoss, VitsGeneratorLoss  # noqa: E402
from python.xvapitch.model import VitsDiscriminator  # noqa: E402

T = 2053      # prime: every period discriminator reflect-pads (xvapitch/hifigan.py:352-355)


def fill_state(params, spec, gen):
    """Seeded weights: directions N(0, 1/fan_in), biases U(-0.05, 0.05), then gains = ||v|| * U(1, 1.1)."""
    with torch.no_grad():
        for k, shape in spec:
            if k.endswith("weight_v"):
                params[k].copy_(torch.randn(shape, generator=gen) / np.sqrt(int(np.prod(shape[1:]))))
            elif not k.endswith("weight_g"):
                params[k].copy_((torch.rand(shape, generator=gen) * 2 - 1) * 0.05)
        for k, shape in spec:
            if k.endswith("weight_g"):
                v = params[k[:-1] + "v"]
                params[k].copy_(v.flatten(1).norm(dim=1).view(shape) * (1.0 + 0.1 * torch.rand(shape, generator=gen)))


def main():
    torch.manual_seed(1234)
    D = VitsDiscriminator()
    gen = torch.Generator().manual_seed(11)
    named = list(D.named_parameters())
    spec = [(k, tuple(p.shape)) for k, p in named]
    fill_state(dict(named), spec, gen)
    x = torch.randn(2, 1, T, generator=gen) * 0.3
    x_hat = (x + 0.2 * torch.randn(2, 1, T, generator=gen)).requires_grad_(True)
    out = {"spec_keys": np.array([k for k, _ in spec]), "spec_shapes": np.array([str(sh) for _, sh in spec]),
           "x": x.numpy(), "x_hat": x_hat.detach().numpy()}
    # discriminator step: scores on (x, x_hat.detach()), LSGAN loss, parameter gradients
    sr, fr, sg, fg = D(x, x_hat.detach())
    loss_d, _, _ = VitsDiscriminatorLoss.discriminator_loss(sr, sg)
    D.zero_grad()
    loss_d.backward()
    for i, (a, b) in enumerate(zip(sr, sg)):
        out[f"score_real/{i}"], out[f"score_fake/{i}"] = a.detach().numpy(), b.detach().numpy()
    out["fmap_real_meanabs"] = np.array([float(f.abs().mean()) for fs in fr for f in fs])   # 7 + 5 * 6
    out["fmap_fake_meanabs"] = np.array([float(f.abs().mean()) for fs in fg for f in fs])
    out["loss_disc"] = np.float64(loss_d.item())
    out["grad_norms"] = np.array([float(p.grad.double().norm()) for _, p in named])
    out["grad/nets.0.convs.1.weight_v"] = dict(named)["nets.0.convs.1.weight_v"].grad.numpy()   # the 4-channel groups
    out["grad/nets.0.convs.0.weight_v"] = dict(named)["nets.0.convs.0.weight_v"].grad.numpy()
    out["grad/nets.0.convs.0.bias"] = dict(named)["nets.0.convs.0.bias"].grad.numpy()
    # generator side: adversarial + feature-matching loss and their gradient wrt the generated waveform. The feature
    # loss is recorded with its arguments by NAME; xvapitch/losses.py:196 passes (fake, real) positionally, which puts
    # the detach on the generated features -- a quirk of that caller, not of the function.
    sr, fr, sg, fg = D(x, x_hat)
    loss_gen, _ = VitsGeneratorLoss.generator_loss(sg)
    loss_feat = VitsGeneratorLoss.feature_loss(feats_real=fr, feats_generated=fg)
    (dw_gen,) = torch.autograd.grad(loss_gen, x_hat, retain_graph=True)
    (dw_feat,) = torch.autograd.grad(loss_feat, x_hat)
    out["loss_gen"], out["loss_feat"] = np.float64(loss_gen.item()), np.float64(loss_feat.item())
    out["dwave_gen"], out["dwave_feat"] = dw_gen.numpy(), dw_feat.numpy()
    np.savez_compressed(os.path.join(HERE, "vits_discriminator.npz"), **out)
    print(len(named), "parameter tensors,", sum(p.numel() for _, p in named), "parameters; losses", loss_d.item(),
          loss_gen.item(), loss_feat.item())


if __name__ == "__main__":
    main()
