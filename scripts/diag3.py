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

orig = ops.gemm_launch
ops.gemm_launch = lambda args, ref=False: orig(args, True)
for Tm in (150, 152, 149):
 for seed in (11, 12):
  for ragged in (False, True):
    stage = 3
    x, y = ofp.synthetic_batch(4, 40, Tm, seed=seed, ragged=ragged)
    sd = ofp.make_state(1234)
    m = fp.FastPitch(device="cuda:0"); m.load_state_dict(sd); m.training_stage = stage; m.train(); m.p_drop = 0.0
    crit = fp.FastPitchLoss(); crit.training_stage = stage
    cx = [t.cuda() if torch.is_tensor(t) else t for t in x]; cy = [t.cuda() if torch.is_tensor(t) else t for t in y]
    out = m(cx); loss, meta = crit(out, cy); m.zero_grad(); m.backward(crit, 1.0); torch.cuda.synchronize()
    keys = ofp.trainable_keys(stage)
    leaves = {k: sd[k].detach().requires_grad_(True) for k in keys}
    work = dict(sd); work.update(leaves)
    o = ofp.forward(work, x, stage)
    total, _ = ofp.loss(o, y, stage)
    grads = dict(zip(keys, torch.autograd.grad(total, [leaves[k] for k in keys], allow_unused=True)))
    got = m.grads(keys)
    errs = sorted(((rel(got[k], grads[k]), k) for k in keys if grads[k] is not None and float(grads[k].norm()) > 0), reverse=True)
    print(f"Tm={Tm} seed={seed} ragged={ragged} mel_lens={x[3].tolist()} in_lens={x[1].tolist()}: " + "; ".join(f"{k.replace('encoder.layers','eL').replace('decoder.layers','dL')}={e:.1e}" for e, k in errs[:5]),
          "| n>1e-5:", sum(e > 1e-5 for e, _ in errs), "mel_out", f"{rel(out[0], o[0]):.1e}")
