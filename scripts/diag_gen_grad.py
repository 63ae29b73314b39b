"""Per-tensor gradient error of the HiFi-GAN generator vs the CPU oracle (diagnostic for tests/test_hifigan_gpu.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from oracle import hifigan as ohg
from xva_trainer_b200 import hifigan as hg
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_hifigan_gpu import _config, _run, rel

for T, seed in [(6, 3), (9, 4), (6, 5), (12, 6)]:
    sd = ohg.make_generator_state(seed, scale=0.7)
    g = hg.Generator(_config(), device="cuda:0")
    g.load_state_dict(sd)
    g.train()
    gen = torch.Generator().manual_seed(seed)
    mel = torch.randn(2, 80, T, generator=gen)
    w = torch.randn(2, 1, 256 * T, generator=gen)
    y, yo, grads, want = _run(g, sd, mel, w)
    num = den = 0.0
    errs = []
    for k, gr in grads.items():
        n_ = float((gr.double() - want[k].double()).pow(2).sum()); d_ = float(want[k].double().pow(2).sum())
        num += n_; den += d_
        errs.append((n_, k, rel(gr, want[k]), d_ ** 0.5))
    errs.sort(reverse=True)
    print(f"T={T} seed={seed}: fwd rel {rel(y, yo):.2e} global grad rel {(num / den) ** 0.5:.3e}")
    for n_, k, e, nm in errs[:8]:
        print(f"   {k:40s} rel {e:.3e} norm {nm:.3e} share of err^2 {n_ / num:.2f}")
