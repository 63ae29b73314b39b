"""CPU-side checks of the xVAPitch tier: size-independent properties of the oracle
(oracle/vits.py, pinned to the reference by tests/test_oracle_golden.py) that the GPU parity tests rely on, and the
product modules' refusal to run without a CUDA device (there is no CPU path)."""
import numpy as np
import pytest
import torch

from oracle import hifigan as ohg
from oracle import vits as ov


def test_maximum_path_is_a_monotonic_surjective_alignment():
    """Every valid frame is assigned to exactly one token, token indices never decrease and advance by at most one per
    frame, the first frame sits on token 0 and the last valid frame on the last valid token; nothing outside the mask."""
    g = torch.Generator().manual_seed(0)
    B, tx, ty = 5, 17, 43
    x_lens, y_lens = np.array([17, 9, 1, 12, 17]), np.array([43, 20, 7, 12, 17])
    value = torch.randn(B, tx, ty, generator=g) * 3
    value[4] = 0.0                                                   # all ties
    path = ov.maximum_path(value, x_lens, y_lens).numpy()
    for b in range(B):
        p = path[b]
        assert p[x_lens[b]:].sum() == 0 and p[:, y_lens[b]:].sum() == 0
        cols = p[:x_lens[b], :y_lens[b]]
        assert np.array_equal(cols.sum(0), np.ones(y_lens[b]))
        tok = cols.argmax(0)
        assert tok[0] == 0 and tok[-1] == x_lens[b] - 1
        assert np.all(np.diff(tok) >= 0) and np.all(np.diff(tok) <= 1)


def test_flow_reverse_inverts_forward_and_preserves_volume():
    """reverse(forward(x)) = x on the valid frames (mean-only coupling: the Jacobian is unit-triangular)."""
    gen = torch.Generator().manual_seed(1)
    sd = {}
    for k, sh in ov.flow_spec():
        if k.endswith("weight_g"):
            continue
        sd[k] = torch.randn(sh, generator=gen) * (0.3 / np.sqrt(max(1, int(np.prod(sh[1:]))))) if len(sh) > 1 else torch.randn(sh, generator=gen) * 0.02
    for k, sh in ov.flow_spec():
        if k.endswith("weight_g"):
            sd[k] = sd[k[:-1] + "v"].flatten(1).norm(dim=1).view(sh)
    x = torch.randn(2, 192, 11, generator=gen)
    mask = ov.sequence_mask([11, 6], 11)[:, None, :].float()
    cond = torch.nn.functional.normalize(torch.randn(2, 512, 1, generator=gen), dim=1)
    with torch.no_grad():
        z = ov.residual_coupling_blocks(sd, x * mask, mask, cond)
        back = ov.residual_coupling_blocks(sd, z, mask, cond, reverse=True)
    torch.testing.assert_close(back, x * mask, rtol=1e-4, atol=1e-5)
    assert float((z * (1 - mask)).abs().max()) == 0.0


def test_segment_rule_and_kl_properties():
    lens = torch.tensor([40, 33, 32])
    for u in (torch.zeros(3), torch.full((3,), 0.999999), torch.tensor([0.5, 0.25, 0.75])):
        s = ov.segment_starts(u, lens)
        assert bool((s >= 0).all()) and bool((s + ov.SEGMENT <= lens).all())
    with pytest.raises(AssertionError):
        ov.segment_starts(torch.zeros(1), torch.tensor([31]))
    # KL of a distribution with itself at its own mean is -1/2 + 1/2 E[eps^2] -> 0 in expectation; exactly -0.5 * C at z = m
    m = torch.randn(2, 8, 5)
    logs = torch.randn(2, 8, 5) * 0.1
    mask = torch.ones(2, 1, 5)
    assert abs(float(ov.kl_loss(m, logs, m, logs, mask)) - (-0.5 * 8)) < 1e-5
    x = torch.randn(2, 1, 4 * 256)
    assert ov.torch_stft_mel(x).shape == (2, 80, 5)


def test_product_modules_refuse_a_cpu_device():
    from xva_trainer_b200 import capi, hifigan as hg, vits

    with pytest.raises((capi.XvaError, RuntimeError, AssertionError)):
        vits.PosteriorEncoder(513, 192, 192, 5, 1, 16, cond_channels=512, device="cpu")
    with pytest.raises((capi.XvaError, RuntimeError, AssertionError)):
        hg.VitsDiscriminator(device="cpu")
    with pytest.raises((capi.XvaError, RuntimeError, AssertionError)):
        vits.ResidualCouplingBlocks(192, 192, 5, 1, 4, cond_channels=512, device="cpu")
