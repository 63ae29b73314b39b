"""Monotonic alignment search on the GPU (xva_mas_width1, csrc/mas.cu) vs the reference's numba b_mas (golden fixture
tests/golden/mas.npz, recorded by running /root/reference's alignment.py) and vs the numpy oracle. Integer index path:
the hard alignment and the durations must be IDENTICAL, not close."""
import os

import numpy as np
import pytest
import torch

from oracle import fastpitch as ofp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["small", "mid", "flat", "short_mel"])
def test_mas_matches_reference_golden(lib, case):
    from xva_trainer_b200 import ops
    g = np.load(os.path.join(GOLD, "mas.npz"))
    attn = torch.from_numpy(g[f"{case}/attn"]).cuda()
    il, ol = torch.from_numpy(g[f"{case}/in_lens"]).cuda(), torch.from_numpy(g[f"{case}/out_lens"]).cuda()
    hard, durs = ops.mas_width1(attn, il, ol)
    want = torch.from_numpy(g[f"{case}/hard"])
    assert torch.equal(hard.cpu(), want)
    assert torch.equal(durs.cpu().long(), want.sum(2)[:, 0, :].long())          # attn_hard.sum(2), model.py:318
    # the same values handed in as fp32 log-probabilities: the recurrence itself is exact
    with np.errstate(divide="ignore"):
        la = np.log(g[f"{case}/attn"])
    hard2, _ = ops.mas_width1(torch.from_numpy(la).cuda(), il, ol, is_log=True)
    assert torch.equal(hard2.cpu(), torch.from_numpy(ofp.b_mas(la, g[f"{case}/in_lens"], g[f"{case}/out_lens"], is_log=True)))


def test_mas_full_size_properties(lib):
    """BASELINE configs[1] size (32 x 880 frames x 160 tokens): exact agreement with the oracle on log-probability input,
    and the structural properties of a monotone alignment (one token per frame, non-decreasing, starts at token 0, ends
    at the last token, durations sum to the frame count)."""
    from xva_trainer_b200 import ops
    B, Tm, Tt = 32, 880, 160
    r = np.random.RandomState(7)
    i = np.arange(Tm)[:, None] / Tm
    j = np.arange(Tt)[None, :] / Tt
    logits = (-3.0 * (j - i) ** 2 * Tt + r.randn(B, 1, Tm, Tt)).astype(np.float32)
    la = (logits - np.log(np.exp(logits).sum(-1, keepdims=True))).astype(np.float32)
    il = r.randint(96, Tt + 1, size=B)
    ol = r.randint(600, Tm + 1, size=B)
    hard, durs = ops.mas_width1(torch.from_numpy(la).cuda(), torch.from_numpy(il).cuda(), torch.from_numpy(ol).cuda(), is_log=True)
    hard, durs = hard.cpu().numpy(), durs.cpu().numpy()
    want = ofp.b_mas(la, il, ol, is_log=True)
    assert np.array_equal(hard, want)
    for b in range(B):
        h = hard[b, 0]
        assert h[ol[b]:].sum() == 0 and h[:, il[b]:].sum() == 0
        assert (h[:ol[b]].sum(1) == 1).all()
        tok = h[:ol[b]].argmax(1)
        assert tok[0] == 0 and tok[-1] == il[b] - 1 and (np.diff(tok) >= 0).all() and (np.diff(tok) <= 1).all()
        assert durs[b].sum() == ol[b] and np.array_equal(durs[b], h.sum(0).astype(np.int32))


@pytest.mark.parametrize("case", ["small", "mid", "ties"])
def test_xvapitch_maximum_path_matches_reference(lib, case):
    """SURVEY 8f rank 1, first piece: xVAPitch's numpy maximum_path (python/xvapitch/util.py:14-53, recorded from the
    unmodified reference by tests/golden/make_golden_xvapitch_mas.py) is xva_mas_width1 with the stay-on-tie flag --
    identical paths, including the fixture made of exact ties."""
    from xva_trainer_b200 import ops
    g = np.load(os.path.join(GOLD, "xvapitch_mas.npz"))
    val = torch.from_numpy(g[f"{case}/value"]).cuda()
    xl, yl = torch.from_numpy(g[f"{case}/x_lens"]).cuda(), torch.from_numpy(g[f"{case}/y_lens"]).cuda()
    path = ops.maximum_path(val, xl, yl)
    assert torch.equal(path.cpu(), torch.from_numpy(g[f"{case}/path"]))
