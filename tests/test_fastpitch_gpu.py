"""FastPitch engine (xva-trainer_b200/fastpitch.py, all math through libxva_b200.so) vs the CPU oracle, which is itself
pinned to the reference's outputs by tests/test_oracle_golden.py, and vs the golden fixtures recorded from the reference.

Two layers of evidence, because the product path multiplies tf32 operands (10-bit mantissa, fp32 accumulate):

  (1) WIRING, exact arithmetic: the same engine with every tap-GEMM routed to the fp32 SIMT checker kernel
      (xva_gemm_ref, same C ABI, same epilogues) must match the oracle to fp32 rounding: forward <= 2e-5, losses
      <= 1e-5, gradients: median <= 1e-5 over tensors and every tensor <= 2e-2. The per-tensor bound is not tighter
      because ReLU gates are discontinuous: one hidden unit whose pre-activation is within fp32 rounding of zero flips
      between the two implementations and moves that layer's weight gradient by ~1e-3..1e-2 (observed sporadically;
      the reference has the same sensitivity to its own cuDNN algorithm choice).
  (2) PRODUCT PATH, tf32 tensor cores (operands rounded to nearest at the producer): the mel output within 1e-3, every
      forward tensor within 2e-3 (dur_pred = exp(x) - 1 near zero: 6e-3), losses within 1e-3 of the oracle; parameter
      gradients within 5e-2 per tensor and 2e-2 on the global gradient vector. Gradient error is dominated by the
      ~0.1 % of ReLU units whose pre-activation is smaller than the tf32 rounding of the conv operands (their gate
      flips), which is inherent to any reduced-precision forward -- the reference's own default path (cuDNN TF32
      convolutions under fp16 autocast, xva_train.py:787) deviates from strict fp32 by more.
  The length-regulator index path is bit-exact in both (dec_lens; regulated rows are exact copies).
Dropout is off for parity (the reference's torch Philox stream cannot be reproduced by a fused kernel).
"""
import os

import numpy as np
import pytest
import torch

from oracle import fastpitch as ofp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# <= 2x the values measured on B200 at this shape (profiles/r02_parity_table.txt, 'toy 4x40x150'): forward tensors
# <= 1.19e-3 (energy_pred), dur_pred 2.6e-3, losses <= 1.2e-4, gradients worst tensor 3.4e-2 (encoder.word_emb: a handful of
# rows), global 9.0e-3. At the BASELINE shape every figure is 3-6x smaller (tests/test_parity_full_gpu.py).
FWD_TOL, LOSS_TOL, GRAD_TOL, GRAD_GLOBAL_TOL = 2e-3, 2.4e-4, 5e-2, 1.8e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _model(lib, sd, stage, training=True):
    from xva_trainer_b200 import fastpitch as fp

    m = fp.FastPitch(device="cuda:0")
    m.load_state_dict(sd)
    m.training_stage = stage
    m.train(training)
    m.p_drop = 0.0
    return fp, m


def _cuda_batch(x, y):
    cx = [t.cuda() if torch.is_tensor(t) else t for t in x]
    cy = [t.cuda() if torch.is_tensor(t) else t for t in y]
    return cx, cy


def test_state_dict_round_trip(lib):
    sd = ofp.make_state(7)
    fp, m = _model(lib, sd, 3)
    out = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in out.items()] == [(k, tuple(s)) for k, s in ofp.state_spec()]
    for k, v in sd.items():
        assert torch.equal(out[k].cpu(), v), k


@pytest.mark.parametrize("stage", [2, 3, 4])
@pytest.mark.parametrize("ragged", [False, True])
def test_step_matches_oracle(lib, stage, ragged):
    B, Tt, Tm = 4, 40, 150
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=11, ragged=ragged)
    sd = ofp.make_state(1234)
    fp, m = _model(lib, {k: v.clone() for k, v in sd.items()}, stage)
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    opt = fp.Lamb(m, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    lr = ofp.noam_lr(50000)
    fp.adjust_learning_rate(50000, opt, 0.1, 1000)
    assert abs(opt.param_groups[0]["lr"] - lr) < 1e-12

    cx, cy = _cuda_batch(x, y)
    out = m(cx)
    loss, meta = crit(out, cy)
    m.zero_grad()
    m.backward(crit, 1.0)
    torch.cuda.synchronize()

    want = ofp.forward(sd, x, stage)
    names = ["mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred", "energy_tgt"]
    for n, g_, w_ in zip(names, out[:8], want[:8]):
        if w_ is None:
            assert g_ is None, n
            continue
        assert g_.shape == w_.shape, (n, g_.shape, w_.shape)
        if w_.dtype == torch.bool:
            assert torch.equal(g_.cpu(), w_), n
        else:
            tol = {"pitch_tgt": 1e-5, "energy_tgt": 1e-5, "mel_out": 1e-3, "dur_pred": 5.2e-3}.get(n, FWD_TOL)
            assert rel(g_, w_) < tol, (n, rel(g_, w_))

    opt_state = {}
    sd_after = {k: v.clone() for k, v in sd.items()}
    wmeta, wgrads = ofp.train_step(sd_after, x, y, stage, lr, opt_state, drop=0.0, training=False)
    for k in ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"):
        a, b = float(meta[k]), float(wmeta[k])
        assert abs(a - b) <= LOSS_TOL * abs(b) + 1e-9, (k, a, b)

    keys = fp.trainable_keys(stage)
    assert keys == ofp.trainable_keys(stage)
    got = m.grads(keys)
    num = den = 0.0
    worst = ("", 0.0)
    for k in keys:
        w_ = wgrads[k]
        if w_ is None:
            assert float(got[k].abs().max()) == 0.0, k
            continue
        assert got[k].shape == w_.shape, k
        e = rel(got[k], w_)
        num += float((got[k].cpu().double() - w_.double()).pow(2).sum())
        den += float(w_.double().pow(2).sum())
        if e > worst[1]:
            worst = (k, e)
        assert e < GRAD_TOL, (k, e)
    assert (num / den) ** 0.5 < GRAD_GLOBAL_TOL, ((num / den) ** 0.5, worst)

    # optimizer: clip_grad_norm_(1000) + LAMB. The first LAMB step is sign-like (m / sqrt(v) = g / |g| * const), so it is
    # compared on IDENTICAL gradients: the oracle's are written into the gradient arena, then the update must agree
    # with lamb.py:40-106 to fp32 rounding.
    A = m.arena
    for k in keys:
        if wgrads[k] is not None:
            A.view(A.g, k).copy_(fp._to_packed(k, wgrads[k]).cuda())
    opt.step()
    torch.cuda.synchronize()
    after = m.state_dict()
    for k in keys:
        delta_w = sd_after[k] - sd[k]
        delta_g = after[k].cpu() - sd[k]
        if float(delta_w.norm()) == 0.0:
            assert float(delta_g.norm()) == 0.0, k
            continue
        assert rel(delta_g, delta_w) < 1e-4, (k, rel(delta_g, delta_w))
        assert rel(after[k], sd_after[k]) < 1e-6, (k, rel(after[k], sd_after[k]))
    for k in after:
        if k not in keys:
            assert torch.equal(after[k].cpu(), sd[k]), f"frozen tensor {k} moved"


@pytest.mark.parametrize("stage", [2, 3, 4])
def test_wiring_exact_with_fp32_checker_gemm(lib, stage, monkeypatch):
    """Evidence layer (1): every contraction through xva_gemm_ref (exact fp32 products) -> fp32-rounding agreement."""
    from xva_trainer_b200 import capi, ops

    orig = ops.gemm_launch
    monkeypatch.setattr(ops, "gemm_launch", lambda args, ref=False: orig(args, True))
    capi.call("xva_set_operand_rounding", 0)
    try:
        _wiring_check(lib, stage)
    finally:
        capi.call("xva_set_operand_rounding", 1)


def _wiring_check(lib, stage):
    B, Tt, Tm = 3, 36, 121
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=21, ragged=True)
    sd = ofp.make_state(4321)
    fp, m = _model(lib, {k: v.clone() for k, v in sd.items()}, stage)
    # exact arithmetic end to end: the fused attention kernels only exist on the tensor cores (their own parity tests are
    # tests/test_attn_fused_gpu.py), so this check runs the unfused chain, whose GEMMs go to the fp32 checker like all others
    m.fused_attn = False
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    cx, cy = _cuda_batch(x, y)
    out = m(cx)
    loss, meta = crit(out, cy)
    m.zero_grad()
    m.backward(crit, 1.0)
    torch.cuda.synchronize()
    want = ofp.forward(sd, x, stage)
    for n, g_, w_ in zip(range(8), out[:8], want[:8]):
        if w_ is None or w_.dtype == torch.bool:
            continue
        assert rel(g_, w_) < 2e-5, (n, rel(g_, w_))
    wmeta, wgrads = ofp.train_step({k: v.clone() for k, v in sd.items()}, x, y, stage, 1e-3, {}, drop=0.0, training=False)
    for k in ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"):
        assert abs(float(meta[k]) - float(wmeta[k])) <= 1e-5 * abs(float(wmeta[k])) + 1e-9, k
    # train_step returns clipped grads; the clip coefficient is 1 here (norm << 1000)
    errs = sorted(rel(g_, wgrads[k]) for k, g_ in m.grads(fp.trainable_keys(stage)).items()
                  if wgrads[k] is not None and float(wgrads[k].norm()) > 0)
    assert errs[len(errs) // 2] < 1e-5, errs[len(errs) // 2]
    assert errs[-1] < 2e-2, errs[-1]


@pytest.mark.parametrize("stage", [2, 3, 4])
def test_forward_matches_reference_golden(lib, stage):
    """Outputs recorded from the unmodified reference modules (tests/golden/make_golden.py) on the same weights."""
    g = np.load(os.path.join(GOLD, "fastpitch_small.npz"))
    t = torch.from_numpy
    B, Tt = g["in/text"].shape
    Tm = g["in/mel"].shape[2]
    x = [t(g["in/text"]), t(g["in/in_lens"]), t(g["in/mel"]), t(g["in/mel_lens"]), t(g["in/pitch"]), t(g["in/energy"]),
         None, None, t(g["in/durs"]), torch.full((B,), float(Tt)), torch.full((B,), float(Tm)), ["synthetic"] * B]
    y = [x[2], x[1], x[3], x[9]]
    fp, m = _model(lib, ofp.make_state(1234), stage)
    crit = fp.FastPitchLoss()
    crit.training_stage = stage
    cx, cy = _cuda_batch(x, y)
    out = m(cx)
    loss, meta = crit(out, cy)
    names = ["mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred", "energy_tgt"]
    seen = 0
    for n, v in zip(names, out[:8]):
        key = f"s{stage}/out/{n}"
        if key in g.files:
            want = t(g[key])
            assert v is not None and tuple(v.shape) == tuple(want.shape), n
            if want.dtype == torch.bool:
                assert torch.equal(v.cpu(), want)
            else:
                tol = {"mel_out": 1e-3, "dur_pred": 6e-3}.get(n, FWD_TOL)
                assert rel(v, want) < tol, (n, rel(v, want))
            seen += 1
    assert seen >= 2
    for k in ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"):
        key = f"s{stage}/loss/{k}"
        if key in g.files:
            # (this fixture's batch is 3 x 14 tokens x 50 frames: few elements per loss term, bound as in round 1)
            assert abs(float(meta[k]) - float(g[key])) <= 1e-3 * abs(float(g[key])) + 1e-9, (k, float(meta[k]), float(g[key]))
    m.zero_grad()
    m.backward(crit, 1.0)
    got = m.grads(fp.trainable_keys(stage))
    for k, gr in got.items():
        nk = f"s{stage}/grad/{k}/norm"
        if nk in g.files:
            want_norm = float(g[nk])
            assert abs(float(gr.double().norm()) - want_norm) <= 3e-2 * want_norm + 1e-9, (k, float(gr.norm()), want_norm)


def test_dropout_training_step_is_finite_and_replays(lib):
    """With dropout on, backward must re-derive exactly the masks forward used: two identical steps from the same
    counter give bit-identical gradients, a different counter gives different ones."""
    B, Tt, Tm = 3, 32, 100
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=5)
    cx, cy = _cuda_batch(x, y)
    fp, m = _model(lib, ofp.make_state(99), 3)
    m.p_drop = 0.1
    crit = fp.FastPitchLoss()

    def run():
        out = m(cx)
        crit(out, cy)
        m.zero_grad()
        m.backward(crit, 1.0)
        return m.arena.g.clone()

    g1, g2 = run(), run()
    assert torch.isfinite(g1).all()
    # weight gradients are accumulated with fp32 atomics (split-z wgrad), so compare to rounding, not bitwise
    assert rel(g2, g1) < 1e-5
    m.step_dropout()
    g3 = run()
    assert rel(g3, g1) > 1e-2


def test_two_stream_backward_gives_the_same_gradients(lib, monkeypatch):
    """XVA_BWD_STREAMS=1 (the default inside a captured graph) issues the FFT-block weight gradients on a side stream; the
    gradient arena must come out the same (to the rounding of the fp32 atomics the split weight gradient already uses)."""
    B, Tt, Tm = 3, 32, 100
    x, y = ofp.synthetic_batch(B, Tt, Tm, seed=5)
    cx, cy = _cuda_batch(x, y)
    grads = []
    for flag in ("0", "1"):
        monkeypatch.setenv("XVA_BWD_STREAMS", flag)
        fp, m = _model(lib, ofp.make_state(99), 3)
        assert m._side_on() == (flag == "1")
        crit = fp.FastPitchLoss()
        for _ in range(2):
            out = m(cx)
            crit(out, cy)
            m.zero_grad()
            m.backward(crit, 1.0)
        torch.cuda.synchronize()
        grads.append(m.arena.g.clone())
    assert rel(grads[1], grads[0]) < 1e-5
