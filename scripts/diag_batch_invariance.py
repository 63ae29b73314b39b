"""Is the forward independent of the batch composition? The same utterances run as one batch of 4 and as two batches of 2
(different padded lengths) must give the same per-utterance tensors; prints the first layer where they do not."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import fastpitch as fp, synthetic
from oracle import fastpitch as ofp

B, Tt, Tm = 4, 40, 150
x, _ = synthetic.fastpitch_batch(B, Tt, Tm, seed=11, ragged=True)
sd = ofp.make_state(1234)
shard = lambda xs, lo, hi: [t[lo:hi].contiguous() if torch.is_tensor(t) else (t[lo:hi] if isinstance(t, list) else t) for t in xs]
dev = lambda xs: [t.cuda() if torch.is_tensor(t) else t for t in xs]


def run(xs):
    m = fp.FastPitch(device="cuda:0")
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m.training_stage = 3
    m.train()
    m.p_drop = 0.0
    out = m(dev(xs))
    c = m._ctx
    t = {"enc_out": c.dec[0].x * 0}  # placeholder
    t = {}
    for i, L in enumerate(c.enc):
        t[f"enc{i}.x"], t[f"enc{i}.qkv"], t[f"enc{i}.vec"], t[f"enc{i}.y1"], t[f"enc{i}.h"] = L.x, L.qkv, L.vec, L.y1, L.h
    for i, L in enumerate(c.dec):
        t[f"dec{i}.x"], t[f"dec{i}.qkv"], t[f"dec{i}.vec"], t[f"dec{i}.y1"], t[f"dec{i}.h"] = L.x, L.qkv, L.vec, L.y1, L.h
    t["dec_out"], t["mel_out"], t["pitch_pred"], t["energy_pred"] = c.dec_out, out[0], out[4], out[6]
    return {k: v.detach().float().cpu() for k, v in t.items()}, x


whole, _ = run(x)
parts = [run(shard(x, 0, 2))[0], run(shard(x, 2, 4))[0]]
lens_txt, lens_mel = x[1], x[3]
print("fused_attn", os.environ.get("XVA_FUSED_ATTN", "1"), "text lens", lens_txt.tolist(), "mel lens", lens_mel.tolist())
for k in whole:
    worst = 0.0
    for b in range(B):
        a = whole[k][b]
        p = parts[b // 2][k][b % 2]
        n = int(lens_txt[b]) if (k.startswith("enc") or k in ("pitch_pred", "energy_pred")) else int(lens_mel[b])
        if a.dim() >= 2 and a.shape[0] >= n and p.shape[0] >= n:
            a, p = a[:n], p[:n]
        elif a.dim() == 2 and a.shape[-1] >= n:       # [1, Tt] predictor outputs
            a, p = a[..., :n], p[..., :n]
        d = float((a - p).abs().max())
        worst = max(worst, d / max(float(a.abs().max()), 1e-30))
    print(f"{k:14s} max |whole - sharded| / max|whole| over valid rows = {worst:.3e}")
