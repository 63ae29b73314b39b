"""Where does a HiFi-GAN v1 training step (B=16 x 8192) spend its device time?  Diagnostic, not a bench.

  python scripts/prof_hifigan.py [B] [table.txt]     per-shape tap-GEMM table of one eagerly launched step (CUDA events)
  XVA_NCU=1 ncu --profile-from-start off ... python scripts/prof_hifigan.py    one step inside cudaProfilerStart/Stop
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from oracle import hifigan as ohg           # synthetic batch generator only
from xva_trainer_b200 import capi, hifigan as hg, ops


class H(dict):
    __getattr__ = dict.__getitem__


def flops(g):
    if g.mode == 2:
        return 2.0 * g.Z * g.R * g.M * g.N * g.taps
    return 2.0 * g.Z * g.R * g.N * g.K * g.taps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    table = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/hifigan_gemm_table.txt"
    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, learning_rate=2e-4, adam_b1=0.8,
          adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
          fmax_for_loss=None)
    G = hg.Generator(h, device="cuda:0"); G.train()
    mpd = hg.MultiPeriodDiscriminator(device="cuda:0"); mpd.train()
    msd = hg.MultiScaleDiscriminator(device="cuda:0"); msd.train()
    step = hg.HiFiGANStep(G, mpd, msd, h)
    x, y, y_mel = (t.cuda() for t in ohg.synthetic_batch(B, 32, seed=1))
    for _ in range(2):
        step.step(x, y, y_mel)
    torch.cuda.synchronize()
    if os.environ.get("XVA_NCU"):
        torch.cuda.profiler.start()
        step.step(x, y, y_mel)
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
    step.step(x, y, y_mel)
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
    print(open(table).read()[:6000])


if __name__ == "__main__":
    main()
