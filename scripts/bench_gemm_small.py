"""Per-shape timing of the short-K / narrow-N tap-GEMM launches of an FFT block (qkv and output projections and their
gradients): they are epilogue- and latency-bound (2.5 ms of the 13.1 ms FastPitch step at 15-25 % of the HBM roofline).
Policy knobs come from the environment (XVA_GEMM_NTILE, XVA_GEMM_PAIR, XVA_GEMM_SEG): one process per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import ops


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    sd = torch.zeros(1, device="cuda", dtype=torch.int64)
    for B, T in ((32, 880), (32, 160)):
        x, vec, qkv = gen(B, T, 384, seed=1), gen(B, T, 64, seed=2), gen(B, T, 192, seed=3)
        wq, bq, wo = gen(1, 192, 384, seed=4, scale=0.05), gen(192, seed=5), gen(1, 384, 64, seed=6, scale=0.1)
        runs = {
            "o_net fwd  K=64  N=384 (+res, drop)": (lambda: ops.conv_fwd(vec, wo, residual=x, drop_p=0.1, seed=5, seed_dev=sd), 2.0 * B * T * 384 * 64, B * T * (64 + 384 + 384) * 4),
            "qkv fwd    K=384 N=192 (+bias)": (lambda: ops.conv_fwd(x, wq, bias=bq, round_out=True), 2.0 * B * T * 192 * 384, B * T * (384 + 192) * 4),
            "qkv dgrad  K=192 N=384 (+res)": (lambda: ops.conv_dgrad(qkv, wq, residual=x), 2.0 * B * T * 192 * 384, B * T * (192 + 384 + 384) * 4),
            "o_net dgrad K=384 N=64": (lambda: ops.conv_dgrad(x, wo, round_out=True), 2.0 * B * T * 384 * 64, B * T * (384 + 64) * 4),
            "qkv wgrad  M=192 N=384": (lambda: ops.conv_wgrad(qkv, x, (0,)), 2.0 * B * T * 192 * 384, B * T * (192 + 384) * 4),
            "o_net wgrad M=384 N=64": (lambda: ops.conv_wgrad(x, vec, (0,)), 2.0 * B * T * 384 * 64, B * T * (384 + 64) * 4),
        }
        for name, (run, flops, bytes_) in runs.items():
            for _ in range(5):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(30):
                run()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / 30
            print(f"{tag:12s} B={B} T={T:3d} {name:38s} {us:7.1f} us  {flops / us / 1e6:6.1f} TF/s  {bytes_ / us / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
