"""Microbenchmark: weight pack / unpack of the HiFi-GAN models (one launch each), XVA_WNPACK_VEC=0|1 (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import bench
from xva_trainer_b200 import hifigan as hg

dev = torch.device("cuda:0")
h = bench._H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
             resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
G = hg.Generator(h, device=dev)
mpd = hg.MultiPeriodDiscriminator(device=dev)
y = torch.randn(2, 8192, device=dev)
mpd(y, y)                      # builds the packer
opt = hg.AdamW(list(G.parameters()) + list(mpd.parameters()))       # parameters move into one flat arena, as in training


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name, pk in (("generator", G._get_packer()), ("mpd", mpd._packer)):
    nparam = sum(m.weight_v.numel() if hasattr(m, "weight_v") else m.weight.numel() for m, *_ in pk.items)
    al = sum(1 for m, *_ in pk.items if (m.weight_v if hasattr(m, "weight_v") else m.weight).data_ptr() % 16 == 0)
    tf = timeit(pk.pack)
    pk.zero_grads()
    tb = timeit(pk.unpack_grads)
    print(f"vec={os.environ.get('XVA_WNPACK_VEC', '1')} {name}: {nparam / 1e6:.1f} M weights, {al}/{len(pk.items)} tensors 16-byte aligned; "
          f"pack {tf:.1f} us = {nparam * 8 / tf / 1e6:.2f} TB/s, unpack {tb:.1f} us = {nparam * 16 / tb / 1e6:.2f} TB/s")
