"""Launches the four dominant tap-GEMM shapes of the FastPitch decoder FFT block (B=32, T=880) a few times each, for
`ncu --set full` (one capture per shape: -k regex:gemm_tc -s <skip> -c 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import ops

B, T = 32, 880
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
x, h = r(B, T, 384), r(B, T, 1536)
w1, w2 = r(3, 1536, 384) * 0.03, r(3, 384, 1536) * 0.02
b1, b2 = r(1536), r(384)
gam, bet = 1 + 0.1 * r(384), 0.1 * r(384)
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
K3 = (-1, 0, 1)
scores = torch.empty(B, T, 896, device="cuda")
wo = r(1, 384, 64) * 0.1
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which.startswith("hg"):
    # HiFi-GAN generator, last stage at B = 16 x 8192: 32 channels, kernel 11 (resblocks.11): forward with the leaky-ReLU
    # epilogue, input gradient with the gate, weight gradient with the split the engine picks
    Bh, Th, Ch = 16, 8192, 32
    K11 = tuple(range(-5, 6))
    xa = ops.round_tf32_(r(Bh, Th, Ch).reshape(-1), torch.empty(Bh * Th * Ch, device="cuda")).view(Bh, Th, Ch) if False else r(Bh, Th, Ch)
    wa, ba = r(11, Ch, Ch) * 0.05, r(Ch)
    gw = torch.zeros(11, Ch, Ch, device="cuda")
    for _ in range(reps):
        if which == "hg32":
            ops.conv_fwd(xa, wa, K11, bias=ba, act_slope=0.1, round_out=True)
        elif which == "hg32d":
            ops.conv_dgrad(xa, wa, K11, gate=xa, gate_slope=0.1, round_out=True)
        elif which == "hg32w":
            ops.conv_wgrad(xa, xa, K11, out=gw, accumulate=True)
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)
for _ in range(reps):
    if which in ("all", "conv1"):
        ops.conv_fwd(x, w1, K3, bias=b1, relu=True)
    if which in ("all", "conv2ln"):
        ops.conv_fwd(h, w2, K3, bias=b2, residual=x, ln=(gam, bet), save_ln=True, lens=lens, drop_p=0.1, seed=1)
    if which in ("all", "dgrad2"):
        ops.conv_dgrad(x, w2, K3, gate=h)
    if which in ("all", "dgrad1"):
        ops.conv_dgrad(h, w1, K3, residual=x)
    if which in ("all", "wgrad1"):
        ops.conv_wgrad(h, x, K3)
    if which in ("qk",):          # attention scores: K = 64, the launch is all epilogue (101 MB written)
        ops.bmm_nt(x[..., :64], x[..., 64:128], alpha=0.125, out=scores[..., :T])
    if which in ("onet",):        # o_net: K = 64, residual + dropout epilogue
        ops.conv_fwd(x[..., :64], wo, residual=x, drop_p=0.1, seed=1)
    if which in ("conv2",):       # ConvFF second conv, un-fused: bias + dropout + residual
        ops.conv_fwd(h, w2, K3, bias=b2, residual=x, drop_p=0.1, seed=1)
torch.cuda.synchronize()
print("done")
