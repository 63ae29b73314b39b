"""HiFi-GAN generator forward + backward at the north-star kernel size (880-frame mels -> 225 280 samples per item):
per-shape table of the tap-GEMM launches (ResBlock1 convs of every stage, transposed-conv phases, conv_pre / conv_post)
with achieved TFLOP/s and the algorithmic HBM bytes of each launch, so each can be put against the roofline that bounds
it (tensor for ch >= 128, HBM for the 32 / 64-channel stages -- SURVEY.md 8d).

    python scripts/bench_generator_large.py [B=8] [frames=880] [table.txt]

B = 8 keeps the saved activations (~12 tensors of B x 225280 x 32 fp32 per ResBlock) near 25 GB; the tile counts are
already in the thousands, so per-launch efficiency does not depend on B beyond that."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import capi, hifigan as hg, ops


class H(dict):
    __getattr__ = dict.__getitem__


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 880
    table = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/generator_large_table.txt"
    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
    G = hg.Generator(h, device="cuda:0")
    G.train()
    g = torch.Generator(device="cuda").manual_seed(0)
    mel = torch.randn(B, 80, frames, device="cuda", generator=g)
    dy = torch.randn(B, 1, frames * 256, device="cuda", generator=g) * 1e-3

    def step():
        G.zero_grad()
        G(mel)
        G.backward(dy)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rec = []
    orig = ops.gemm_launch

    def timed(a, ref=False):
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        orig(a, ref)
        y.record()
        if a.mode == 2:
            fl = 2.0 * a.Z * a.R * a.M * a.N * a.taps
            by = 4.0 * a.Z * a.R * (a.M + a.N)                  # dy + x read once, dw negligible
        else:
            fl = 2.0 * a.Z * a.R * a.N * a.K * a.taps
            by = 4.0 * a.Z * a.R * (a.K + a.N * (1 + (1 if a.residual else 0) + (1 if a.gate else 0) + (1 if a.out_act else 0)))
        rec.append((x, y, fl, by, (a.mode, a.Z, a.R, a.M, a.N, a.K, a.taps, a.flags)))

    ops.gemm_launch = timed
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    ops.gemm_launch = orig
    agg = {}
    for x, y, fl, by, shape in rec:
        e = agg.setdefault(shape, [0, 0.0, 0.0, 0.0])
        e[0] += 1; e[1] += x.elapsed_time(y); e[2] += fl; e[3] += by
    tot_ms = sum(v[1] for v in agg.values()); tot_f = sum(v[2] for v in agg.values())
    with open(table, "w") as fh:
        fh.write(f"# HiFi-GAN generator forward + backward, B={B} x {frames} frames ({frames * 256} samples/item): {ms:.1f} ms/step eager "
                 f"= {B * frames * 256 / ms / 1e3:.2f} M samples/s; {len(rec)} tap-GEMM launches {tot_ms:.1f} ms, {tot_f / 1e12:.2f} TFLOP "
                 f"= {tot_f / tot_ms / 1e9:.0f} TFLOP/s; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB\n")
        fh.write("mode Z R M N K taps flags | launches ms GFLOP TFLOP/s algorithmic-GB/s\n")
        for shape, (n, t, f, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(" ".join(str(v) for v in shape) + f" | {n} {t:.3f} {f / 1e9:.1f} {f / t / 1e9:.1f} {by / t / 1e6:.0f}\n")
    print(open(table).read()[:5000])


if __name__ == "__main__":
    main()
