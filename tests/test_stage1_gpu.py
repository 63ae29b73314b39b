"""FastPitch training stage 1 (the aligner) on the B200 engine vs the CPU oracle and the goldens recorded from the
reference (tests/golden/make_golden_stage1.py): csrc/align.cu kernels one by one through the C ABI, then the whole
stage-1 step (ConvAttention on the tap-GEMM, score kernel, MAS, CTC + binarization losses, backward, LAMB).

Same two layers of evidence as tests/test_fastpitch_gpu.py: exact wiring with every GEMM routed to the fp32 checker,
and the tf32 product path within the tolerances stated there. The index path (hard alignment, durations) is compared
exactly: MAS on the engine's own soft attention must equal the oracle's MAS on those same values everywhere, and must
equal the oracle's end-to-end hard alignment wherever the two soft attentions do not differ by a near-tie.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fastpitch as ofp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _score_reference(q, k, prior, in_lens):
    """attention.py:203-219 in fp64 torch. q [B,Tm,C], k [B,Tt,C] -> (logprob, soft) [B,Tm,Tt]."""
    sq = (q.unsqueeze(2) - k.unsqueeze(1)).pow(2).sum(-1)
    lp = torch.log_softmax(-0.0005 * sq, dim=2) + torch.log(prior + 1e-8)
    pad = torch.arange(k.shape[1])[None, :] >= in_lens[:, None]
    soft = torch.softmax(lp.masked_fill(pad[:, None, :], float("-inf")), dim=2)
    return lp, soft


@pytest.mark.parametrize("B,Tm,Tt,scale", [(3, 50, 14, 1.0), (2, 131, 45, 6.0), (2, 70, 160, 3.0)])
def test_attn_score_kernels(lib, B, Tm, Tt, scale):
    from xva_trainer_b200 import ops

    g = torch.Generator().manual_seed(Tm)
    C = 80
    q = (torch.randn(B, Tm, C, generator=g) * scale).double().requires_grad_(True)
    k = (torch.randn(B, Tt, C, generator=g) * scale).double().requires_grad_(True)
    prior = torch.rand(B, Tm, Tt, generator=g).double()
    prior[0, 0, 0] = 0.0                                   # log(0 + 1e-8)
    in_lens = torch.randint(max(1, Tt // 2), Tt + 1, (B,), generator=g)
    in_lens[0] = Tt
    want_lp, want_soft = _score_reference(q, k, prior, in_lens)

    qb = torch.zeros(B, Tm, 96, device="cuda")
    kb = torch.zeros(B, Tt, 96, device="cuda")
    qb[..., :C] = q.detach().float().cuda()
    kb[..., :C] = k.detach().float().cuda()
    lens32 = in_lens.to(torch.int32).cuda()
    lp, soft, pr = ops.attn_score_fwd(qb[..., :C], kb[..., :C], prior.float().cuda(), lens32)
    torch.cuda.synchronize()
    assert rel(lp, want_lp) < 1e-5 and rel(soft, want_soft) < 1e-5, (rel(lp, want_lp), rel(soft, want_soft))
    for b in range(B):
        assert float(soft[b, :, int(in_lens[b]):].abs().max() if int(in_lens[b]) < Tt else 0.0) == 0.0

    # backward of an arbitrary linear functional of logprob: sum(logprob * G)
    G = torch.randn(B, Tm, Tt, generator=g).double()
    (want_lp * G).sum().backward()
    dq, dk = ops.attn_score_bwd(G.float().cuda(), lp, pr, qb[..., :C], kb[..., :C])
    torch.cuda.synchronize()
    # dq / dk are rounded to tf32 on store (they feed the projection stacks' gradient GEMMs): 2^-11 relative
    assert rel(dq, q.grad) < 6e-4 and rel(dk, k.grad) < 6e-4, (rel(dq, q.grad), rel(dk, k.grad))
    assert dq.stride(1) == 96 and float(dq.as_strided((B, Tm, 16), (Tm * 96, 96, 1), 80).abs().max()) == 0.0


@pytest.mark.parametrize("B,Tm,Tt", [(3, 23, 8), (4, 150, 37), (2, 300, 160)])
def test_attn_ctc_kernel(lib, B, Tm, Tt):
    from xva_trainer_b200 import ops

    g = torch.Generator().manual_seed(Tt)
    lp = (torch.randn(B, 1, Tm, Tt, generator=g) * 2).log_softmax(-1)
    in_lens = torch.randint(max(1, Tt // 2), Tt + 1, (B,), generator=g)
    out_lens = torch.maximum(torch.randint(Tm // 2, Tm + 1, (B,), generator=g), in_lens)
    in_lens[0], out_lens[0] = Tt, Tm
    if B > 2:
        out_lens[1] = in_lens[1]                           # exactly one frame per token: a single admissible path
        out_lens[2] = max(1, int(in_lens[2]) - 2)          # fewer frames than tokens: impossible -> zero_infinity
    ref = lp.double().requires_grad_(True)
    want = ofp.attention_ctc_loss(ref, in_lens, out_lens)
    want.backward()
    cost, grad = ops.attn_ctc(lp[:, 0].contiguous().cuda(), in_lens.to(torch.int32).cuda(), out_lens.to(torch.int32).cuda())
    torch.cuda.synchronize()
    assert abs(float(cost.mean()) - float(want)) <= 2e-6 * abs(float(want)), (float(cost.mean()), float(want))
    if B > 2:
        assert float(cost[2]) == 0.0 and float(grad[2].abs().max()) == 0.0
    assert rel(grad, ref.grad[:, 0]) < 2e-5, rel(grad, ref.grad[:, 0])
    # and against the written-out fp64 recursion of the oracle, utterance by utterance
    for b in range(B):
        c, _ = ofp.ctc_recursion(lp[b, 0].numpy(), int(in_lens[b]), int(out_lens[b]))
        assert abs(float(cost[b]) - c) <= 2e-6 * abs(c) + 1e-12, (b, float(cost[b]), c)


def test_attn_binarization_loss_and_gradient(lib):
    from xva_trainer_b200 import fastpitch as fp, ops

    g = torch.Generator().manual_seed(3)
    B, Tm, Tt = 3, 40, 21
    in_lens, out_lens = torch.tensor([21, 15, 9]), torch.tensor([40, 33, 20])
    logits = torch.randn(B, 1, Tm, Tt, generator=g).double() * 3
    pad = torch.arange(Tt)[None, :] >= in_lens[:, None]
    logits = logits.requires_grad_(True)
    soft = torch.softmax(logits.masked_fill(pad[:, None, None, :], float("-inf")), dim=3)
    hard = torch.from_numpy(ofp.b_mas(soft.detach().float().numpy(), in_lens.numpy(), out_lens.numpy())).double()
    gctc = torch.randn(B, 1, Tm, Tt, generator=g).double()
    want = ofp.attention_binarization_loss(hard, soft)
    (0.7 * want + 1.3 * (logits * gctc).sum()).backward()
    crit = fp.AttentionBinarizationLoss()
    got = crit(hard.float().cuda(), soft.detach().float().cuda())
    assert abs(float(got) - float(want)) <= 1e-6 * abs(float(want))
    gg = ops.attn_grad_combine(gctc.float().cuda().contiguous(), 1.3, *crit.saved(), bw=0.7)
    torch.cuda.synchronize()
    assert rel(gg, logits.grad) < 1e-6, rel(gg, logits.grad)
    g0 = ops.attn_grad_combine(gctc.float().cuda().contiguous(), 0.5)
    assert rel(g0, 0.5 * gctc) < 1e-7


def _model(sd, training=True):
    from xva_trainer_b200 import fastpitch as fp

    m = fp.FastPitch(device="cuda:0")
    m.load_state_dict(sd)
    m.training_stage = 1
    m.train(training)
    m.p_drop = 0.0
    return fp, m


def _cuda(x):
    return [t.cuda() if torch.is_tensor(t) else t for t in x]


def _run_step(sd, x, y, klw):
    fp, m = _model({k: v.clone() for k, v in sd.items()})
    crit = fp.FastPitchLoss()
    crit.training_stage = 1
    kl = fp.AttentionBinarizationLoss()
    out = m(_cuda(x))
    loss, meta = crit(out, _cuda(y))
    klv = kl(out[9], out[8])
    m.zero_grad()
    m.backward(crit, 1.0, kl=(kl, klw))
    torch.cuda.synchronize()
    return fp, m, out, meta, klv


def _check_step(sd, x, y, klw, fwd_tol, loss_tol, grad_tol, grad_median_tol=None):
    fp, m, out, meta, klv = _run_step(sd, x, y, klw)
    want = ofp.forward(sd, x, 1)
    for i in range(8):
        assert out[i] is None
    assert tuple(out[8].shape) == tuple(want[8].shape) and tuple(out[11].shape) == tuple(want[11].shape)
    assert rel(out[8], want[8]) < fwd_tol and rel(out[11], want[11]) < fwd_tol, (rel(out[8], want[8]), rel(out[11], want[11]))
    # index path: MAS of the engine's soft attention == the oracle's MAS on the same values, bit for bit
    same_in = ofp.b_mas(out[8].cpu().numpy(), x[1].numpy(), x[3].numpy())
    assert np.array_equal(out[9].cpu().numpy(), same_in)
    assert torch.equal(out[10].cpu(), torch.from_numpy(same_in).sum(2)[:, 0, :])
    agree = float((out[9].cpu() == want[9]).float().mean())
    assert agree > 0.995, agree
    wtotal, wmeta = ofp.loss(want, y, 1, kl_weight=klw)
    assert abs(float(meta["attn_loss"]) - float(wmeta["attn_loss"])) <= loss_tol * abs(float(wmeta["attn_loss"])), \
        (float(meta["attn_loss"]), float(wmeta["attn_loss"]))
    if agree == 1.0:
        assert abs(klw * float(klv) - float(wmeta["kl_loss"])) <= loss_tol * abs(float(wmeta["kl_loss"])) + 1e-9
    _, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, 1, 1e-3, {}, drop=0.0, training=False,
                               kl_weight=klw, clip=1e9)
    keys = fp.trainable_keys(1)
    assert keys == ofp.trainable_keys(1)
    got = m.grads()
    errs = []
    for k, gr in got.items():
        if k not in keys:
            assert float(gr.abs().max()) == 0.0, f"{k}: gradient outside the stage-1 set"
            continue
        assert gr.shape == wgrads[k].shape, k
        e = rel(gr, wgrads[k])
        errs.append(e)
        assert e < grad_tol, (k, e)
    if grad_median_tol is not None:
        assert sorted(errs)[len(errs) // 2] < grad_median_tol, sorted(errs)
    return fp, m, wgrads


@pytest.mark.parametrize("mel_scale", [1.0, 10.0])
def test_stage1_step_matches_oracle(lib, mel_scale):
    x, y = ofp.synthetic_batch(4, 40, 150, seed=13, ragged=True, prior=True)
    x[2] = x[2] * mel_scale
    y[0] = x[2]
    sd = ofp.make_state(1234)
    fp, m, wgrads = _check_step(sd, x, y, 0.5, fwd_tol=2e-3, loss_tol=1e-3, grad_tol=5e-2)
    # optimizer on identical gradients (as in test_fastpitch_gpu): only the 11 stage-1 tensors move
    A = m.arena
    keys = fp.trainable_keys(1)
    for k in keys:
        A.view(A.g, k).copy_(fp._to_packed(k, wgrads[k]).cuda())
    opt = fp.Lamb(m, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    lr = ofp.noam_lr(50000)
    fp.adjust_learning_rate(50000, opt, 0.1, 1000)
    opt.step()
    torch.cuda.synchronize()
    sd_after = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ofp.lamb_step(sd_after, wgrads, {}, lr)
    after = m.state_dict()
    for k in after:
        if k in keys:
            assert rel(after[k], sd_after[k]) < 1e-6, (k, rel(after[k], sd_after[k]))
            assert rel(after[k].cpu() - sd[k], sd_after[k] - sd[k]) < 1e-4, k
        else:
            assert torch.equal(after[k].cpu(), sd[k]), f"{k} moved in stage 1 (the reference leaves grad = None)"


def test_stage1_wiring_exact_with_fp32_checker_gemm(lib, monkeypatch):
    """Every contraction through xva_gemm_ref (exact fp32 products), operand rounding off -> fp32-rounding agreement."""
    from xva_trainer_b200 import capi, ops

    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        x, y = ofp.synthetic_batch(3, 36, 121, seed=21, ragged=True, prior=True)
        x[2] = x[2] * 4.0
        y[0] = x[2]
        _check_step(ofp.make_state(4321), x, y, 0.5, fwd_tol=2e-5, loss_tol=1e-5, grad_tol=2e-2, grad_median_tol=2e-5)
    finally:
        capi.call("xva_set_operand_rounding", 1)


@pytest.mark.parametrize("case", ["small", "sharp", "short"])
def test_stage1_matches_reference_golden(lib, case):
    """Outputs of the unmodified reference aligner (tests/golden/make_golden_stage1.py) on the same weights."""
    from test_oracle_golden import stage1_batch

    g = np.load(os.path.join(GOLD, "stage1.npz"))
    x, y = stage1_batch(g, case)
    klw = float(g[f"{case}/kl_weight"])
    fp, m, out, meta, klv = _run_step(ofp.make_state(1234), x, y, klw)
    assert rel(out[8], torch.from_numpy(g[f"{case}/attn_soft"])) < 2e-3
    assert rel(out[11], torch.from_numpy(g[f"{case}/attn_logprob"])) < 2e-3
    want_hard = torch.from_numpy(g[f"{case}/attn_hard"])
    agree = float((out[9].cpu() == want_hard).float().mean())
    assert agree > 0.99, agree
    if agree == 1.0:
        assert torch.equal(out[10].cpu(), torch.from_numpy(g[f"{case}/durs"]))
        assert abs(float(klv) - float(g[f"{case}/kl"])) <= 1e-3 * float(g[f"{case}/kl"])
    assert abs(float(meta["attn_loss"]) - float(g[f"{case}/ctc"])) <= 1e-3 * float(g[f"{case}/ctc"])
    got = m.grads(fp.trainable_keys(1))
    for k, gr in got.items():
        want_norm = float(g[f"{case}/grad/{k}/norm"])
        assert abs(float(gr.double().norm()) - want_norm) <= 3e-2 * want_norm + 1e-12, (k, float(gr.norm()), want_norm)


def test_stage1_full_size_properties(lib):
    """BASELINE.json's FastPitch configuration (32 x 880 frames x 160 tokens), where the oracle's [B, Tm, Tt, C]
    broadcast is too large to run in seconds: size-independent properties instead. Rows of attn_soft sum to one over the
    real keys, the hard alignment is monotonic with one mark per frame and sums to the mel length, exp(attn_logprob -
    log prior) is a distribution, the CTC cost is finite and positive, gradients are finite, and a second identical
    step reproduces the first to rounding."""
    from xva_trainer_b200 import fastpitch as fp

    x, y = ofp.synthetic_batch(32, 160, 880, seed=3, ragged=True, prior=True)
    fp, m = _model(ofp.make_state(1234))
    crit = fp.FastPitchLoss()
    crit.training_stage = 1
    kl = fp.AttentionBinarizationLoss()
    cx, cy = _cuda(x), _cuda(y)

    def step():
        out = m(cx)
        loss, meta = crit(out, cy)
        klv = kl(out[9], out[8])
        m.zero_grad()
        m.backward(crit, 1.0, kl=(kl, 1.0))
        return out, float(loss), float(klv), m.arena.g.clone()

    out, loss, klv, g1 = step()
    soft, hard, durs, logprob = out[8][:, 0], out[9][:, 0], out[10], out[11][:, 0]
    in_lens, mel_lens = cx[1], cx[3]
    assert torch.allclose(soft.sum(-1), torch.ones_like(soft.sum(-1)), atol=1e-5)
    sm = torch.exp(logprob - torch.log(cx[7] + 1e-8))
    assert torch.allclose(sm.sum(-1), torch.ones_like(sm.sum(-1)), atol=1e-4)
    assert torch.equal(durs.sum(1).long(), mel_lens)
    for b in (0, 7, 31):
        n, t = int(in_lens[b]), int(mel_lens[b])
        h = hard[b, :t, :n]
        assert torch.equal(h.sum(1), torch.ones(t, device=h.device))
        pos = h.argmax(1)
        assert int(pos[0]) == 0 and int(pos[-1]) == n - 1 and bool(((pos[1:] - pos[:-1]) >= 0).all()) \
            and bool(((pos[1:] - pos[:-1]) <= 1).all())
        assert float(hard[b, t:].abs().max() if t < hard.shape[1] else 0.0) == 0.0
    assert np.isfinite(loss) and loss > 0 and np.isfinite(klv) and klv > 0
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    _, loss2, klv2, g2 = step()
    assert abs(loss2 - loss) <= 1e-6 * abs(loss) and rel(g2, g1) < 1e-5
