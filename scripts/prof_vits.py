"""Where does an xVAPitch --hifi_only training step (batch 16 x 256 frames) spend its device time?  Diagnostic, not a bench.

  python scripts/prof_vits.py [table.txt]     per-shape tap-GEMM table of one eagerly launched step (CUDA events)
  XVA_NCU=1 ncu --profile-from-start off ... python scripts/prof_vits.py    one step inside cudaProfilerStart/Stop
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import bench
from xva_trainer_b200 import capi, hifigan as hg, ops, vits


def flops(g):
    if g.mode == 2:
        return 2.0 * g.Z * g.R * g.M * g.N * g.taps
    return 2.0 * g.Z * g.R * g.N * g.K * g.taps


def main():
    table = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/vits_gemm_table.txt"
    dev = torch.device("cuda:0")
    enc = vits.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device=dev)
    dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2], inference_padding=0,
                              cond_channels=512, conv_pre_weight_norm=False, conv_post_weight_norm=False,
                              conv_post_bias=False, device=dev)
    disc = hg.VitsDiscriminator(device=dev)
    for m in (enc, dec, disc):
        m.train()
    step = vits.HifiOnlyStep(enc, dec, disc)
    lin, lens, wav, dv = (t.to(dev) for t in bench.synthetic_vits_batch(16, 256, 1))
    for _ in range(2):
        step.step(lin, lens, wav, dv)
    torch.cuda.synchronize()
    if os.environ.get("XVA_NCU"):
        torch.cuda.profiler.start()
        step.step(lin, lens, wav, dv)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    rec = []
    orig = ops.gemm_launch

    def timed(g, ref=False):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(g, ref)
        b.record()
        rec.append((a, b, flops(g), (g.mode, g.Z, g.R, g.M, g.N, g.K, g.taps, g.flags, g.split)))

    ops.gemm_launch = timed
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(0.5 * 1.9e9))
    capi.reset_launch_count()
    s0.record()
    step.step(lin, lens, wav, dv)
    s1.record()
    torch.cuda.synchronize()
    ops.gemm_launch = orig
    agg = {}
    for a, b, f, shape in rec:
        e = agg.setdefault(shape, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += a.elapsed_time(b)
        e[2] += f
    tot_ms = sum(v[1] for v in agg.values())
    tot_f = sum(v[2] for v in agg.values())
    with open(table, "w") as fh:
        fh.write(f"# eager step {s0.elapsed_time(s1):.2f} ms device span (GPU parked 0.5 s first); {len(rec)} tap-GEMM launches "
                 f"{tot_ms:.2f} ms, {tot_f / 1e9:.0f} GFLOP; {capi.launch_count()} C-ABI launches\n")
        fh.write("mode Z R M N K taps flags split | launches ms GFLOP TFLOP/s us/launch\n")
        for shape, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(" ".join(str(v) for v in shape) + f" | {n} {t:.3f} {f / 1e9:.2f} {f / (t * 1e-3) / 1e12:.1f} {1e3 * t / n:.1f}\n")
    print(open(table).read()[:5000])


if __name__ == "__main__":
    main()
