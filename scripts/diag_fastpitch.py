"""Diagnostic (not a test): per-tensor error of the FastPitch engine vs the oracle, with the tcgen05 tap-GEMM and with
the exact-fp32 SIMT checker substituted for it, to separate tf32 rounding from wiring bugs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from oracle import fastpitch as ofp
from xva_trainer_b200 import fastpitch as fp, ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def run(stage, use_ref, B=4, Tt=40, Tm=150, ragged=True):
    orig = ops.gemm_launch
    if use_ref:
        ops.gemm_launch = lambda args, ref=False: orig(args, True)
    try:
        x, y = ofp.synthetic_batch(B, Tt, Tm, seed=11, ragged=ragged)
        sd = ofp.make_state(1234)
        m = fp.FastPitch(device="cuda:0"); m.load_state_dict(sd); m.training_stage = stage; m.train(); m.p_drop = 0.0
        crit = fp.FastPitchLoss(); crit.training_stage = stage
        cx = [t.cuda() if torch.is_tensor(t) else t for t in x]; cy = [t.cuda() if torch.is_tensor(t) else t for t in y]
        out = m(cx); loss, meta = crit(out, cy); m.zero_grad(); m.backward(crit, 1.0); torch.cuda.synchronize()
        want = ofp.forward(sd, x, stage)
        names = ["mel_out", "dec_mask", "dur_pred", "log_dur_pred", "pitch_pred", "pitch_tgt", "energy_pred", "energy_tgt"]
        for n, g_, w_ in zip(names, out[:8], want[:8]):
            if w_ is not None and w_.dtype != torch.bool:
                print(f"  stage {stage} ref={use_ref} out {n:14s} rel {rel(g_, w_):.2e}  max|w| {w_.abs().max():.3f}")
        sd2 = {k: v.clone() for k, v in sd.items()}
        wmeta, wgrads = ofp.train_step(sd2, x, y, stage, 1e-3, {}, drop=0.0, training=False)
        for k in meta:
            print(f"  loss {k:26s} got {float(meta[k]):.6f} want {float(wmeta[k]):.6f}")
        got = m.grads(fp.trainable_keys(stage))
        errs = sorted(((rel(got[k], wgrads[k]), k) for k in got if wgrads[k] is not None), reverse=True)
        for e, k in errs[:12]:
            print(f"  grad {k:50s} rel {e:.2e}")
        print("  median grad err", errs[len(errs) // 2][0])
    finally:
        ops.gemm_launch = orig


for stage in (2, 3, 4):
    for use_ref in (True, False):
        run(stage, use_ref)
