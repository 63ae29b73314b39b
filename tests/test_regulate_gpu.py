"""Length regulator / average_pitch kernels vs the oracle and the golden vectors recorded from the reference.
Index path: bit-exact (torch.equal)."""
import os

import numpy as np
import pytest
import torch

from oracle import fastpitch as ofp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ops(lib):
    from xva_trainer_b200 import ops

    return ops


def _run(ops, durs, enc, pace, mx):
    cum, dec = ops.duration_scan(durs.cuda(), pace, mx)
    t_out = int(dec.max().item())
    out, idx = ops.regulate_gather(enc.cuda(), cum, t_out, want_idx=True)
    return out.cpu(), dec.cpu().long(), idx.cpu().long(), cum


@pytest.mark.parametrize("case", ["plain", "pace", "trunc", "pace_trunc"])
def test_regulate_len_golden_bit_exact(lib, case):
    ops = _ops(lib)
    g = np.load(os.path.join(GOLD, "regulate_len.npz"))
    durs, enc = torch.from_numpy(g["durs"]), torch.from_numpy(g["enc"])
    pace, mx = float(g[f"{case}/pace"]), int(g[f"{case}/mel_max_len"])
    out, dec, idx, _ = _run(ops, durs, enc, pace, None if mx < 0 else mx)
    assert torch.equal(dec, torch.from_numpy(g[f"{case}/dec_lens"]))
    assert torch.equal(out, torch.from_numpy(g[f"{case}/enc_rep"]))


@pytest.mark.parametrize("B,Tt,C,maxd,seed", [(1, 1, 4, 3, 0), (2, 7, 8, 1, 1), (32, 160, 384, 12, 2), (5, 300, 64, 40, 3),
                                              (3, 1000, 16, 3, 4)])
def test_regulate_len_random_vs_oracle(lib, B, Tt, C, maxd, seed):
    ops = _ops(lib)
    g = torch.Generator().manual_seed(seed)
    durs = torch.randint(0, maxd + 1, (B, Tt), generator=g).float()
    durs[0, 0] = max(1.0, float(durs[0, 0]))
    enc = torch.randn(B, Tt, C, generator=g)
    for pace, mx in ((1.0, None), (0.9, None), (1.0, max(1, int(durs.sum(1).max()) // 2))):
        want, want_lens = ofp.regulate_len(durs, enc, pace, mx)
        want_idx, _ = ofp.regulate_indices(durs, pace, mx)
        out, dec, idx, cum = _run(ops, durs, enc, pace, mx)
        assert torch.equal(dec, want_lens)
        assert torch.equal(idx[:, :want_idx.shape[1]], want_idx)
        assert torch.equal(out[:, :want.shape[1]], want)
        # backward: contiguous-segment sums == autograd of the oracle
        e = enc.clone().requires_grad_(True)
        w, _ = ofp.regulate_len(durs, e, pace, mx)
        dout = torch.randn(w.shape, generator=g)
        w.backward(dout)
        denc = ops.regulate_scatter(dout.cuda(), cum, Tt).cpu()
        torch.testing.assert_close(denc, e.grad, rtol=1e-5, atol=1e-5)


def test_regulate_len_full_size_roundtrip_property(lib):
    """BASELINE size: every frame row equals the encoder row its index names; rows past dec_len are zero;
    sum of per-token frame counts == dec_len."""
    ops = _ops(lib)
    x, _ = ofp.synthetic_batch(32, 160, 880, seed=5)
    durs = x[8].cuda()
    enc = torch.randn(32, 160, 384, device="cuda")
    cum, dec = ops.duration_scan(durs)
    out, idx = ops.regulate_gather(enc, cum, 880, want_idx=True)
    assert torch.equal(dec.long(), durs.sum(1).long())
    valid = idx >= 0
    assert torch.equal(valid.sum(1), dec.long())
    rows = torch.gather(enc, 1, idx.clamp_min(0).long().unsqueeze(-1).expand(-1, -1, 384))
    assert torch.equal(out, rows * valid.unsqueeze(-1))
    counts = torch.zeros(32, 160, device="cuda").scatter_add_(1, idx.clamp_min(0).long(), valid.float())
    assert torch.equal(counts, durs)


def test_average_pitch_vs_oracle_and_golden(lib):
    ops = _ops(lib)
    g = np.load(os.path.join(GOLD, "regulate_len.npz"))
    out = ops.average_pitch(torch.from_numpy(g["avg/pitch"]).cuda(), torch.from_numpy(g["avg/durs"]).cuda()).cpu()
    torch.testing.assert_close(out, torch.from_numpy(g["avg/out"]), rtol=0, atol=1e-5)
    x, _ = ofp.synthetic_batch(32, 160, 880, seed=6)
    want = ofp.average_pitch(x[4], x[8])
    got = ops.average_pitch(x[4].cuda(), x[8].cuda()).cpu()
    assert torch.equal(got == 0, want == 0)
    torch.testing.assert_close(got, want, rtol=0, atol=1e-5)
