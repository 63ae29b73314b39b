"""Kernel timeline of ONE graph replay of the HiFi-GAN (or xVAPitch --hifi_only) training step, from CUPTI through
torch.profiler: how much of the step runs with 1, 2, 3 ... kernels in flight, and which kernels make up the serial part.
Diagnostic, not a bench (the profiler adds overhead: compare shares, not absolute times).

    python scripts/timeline_hifigan.py [hifigan|vits] [out.txt]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

ge.build()
import bench
from xva_trainer_b200 import graph, hifigan as hg, synthetic, vits


def build(which, dev):
    if which == "fastpitch":
        from xva_trainer_b200 import fastpitch as fp

        model = fp.FastPitch(device=dev, seed=1234)
        model.training_stage = 3
        model.train()
        crit = fp.FastPitchLoss()
        crit.training_stage = 3
        opt = fp.Lamb(model, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        x_cpu, _ = synthetic.fastpitch_batch(32, 160, 880, seed=1234, ragged=False)
        host_lens = (880, int(x_cpu[3].max()))
        x_dev = [t.to(dev) if torch.is_tensor(t) else t for t in x_cpu]
        idx = [i for i, t in enumerate(x_dev) if torch.is_tensor(t)]
        opt.lr_on_device = True

        def step(*tensors):
            xs = list(x_dev)
            for i, t in zip(idx, tensors):
                xs[i] = t
            model.zero_grad()
            out = model(xs, host_lens=host_lens)
            loss, _ = crit(out, [xs[2], xs[1], xs[3], xs[9]])
            model.backward(crit, 1.0)
            opt.step()
            model.step_dropout()
            return loss

        return graph.GraphedStep(step, [x_dev[i] for i in idx], warmup=3)
    if which == "vits":
        enc = vits.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device=dev)
        dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                                  inference_padding=0, cond_channels=512, conv_pre_weight_norm=False,
                                  conv_post_weight_norm=False, conv_post_bias=False, device=dev)
        disc = hg.VitsDiscriminator(device=dev)
        for m in (enc, dec, disc):
            m.train()
        st = vits.HifiOnlyStep(enc, dec, disc)
        inputs = [t.to(dev) for t in bench.synthetic_vits_batch(16, 256, 1)]
        fn = lambda a, b, c, d: st.step(a, b, c, d)
    else:
        h = bench._H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
                     resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, learning_rate=2e-4, adam_b1=0.8,
                     adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
                     fmax_for_loss=None)
        G = hg.Generator(h, device=dev); G.train()
        mpd = hg.MultiPeriodDiscriminator(device=dev); mpd.train()
        msd = hg.MultiScaleDiscriminator(device=dev); msd.train()
        st = hg.HiFiGANStep(G, mpd, msd, h)
        inputs = list(synthetic.hifigan_batch(16, 32, dev, seed=1))
        fn = lambda a, b, c: st.step(a, b, c)
    st.optim_g.lr_on_device = st.optim_d.lr_on_device = True
    return graph.GraphedStep(fn, inputs, warmup=3)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "hifigan"
    out = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/timeline_{which}.txt"
    dev = torch.device("cuda:0")
    gs = build(which, dev)
    for _ in range(5):
        gs()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gs(); e1.record(); torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1)
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        gs()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ker = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()))
    if not ker:
        print("no kernel events recorded"); return
    t0, t1 = ker[0][0], max(k[1] for k in ker)
    span = (t1 - t0) / 1e3
    # sweep: time spent at each concurrency level, and per-kernel-name time while it was the only kernel running
    pts = []
    for i, (s, e, n) in enumerate(ker):
        pts.append((s, 1, i)); pts.append((e, -1, i))
    pts.sort()
    live, last, level_time, solo = set(), t0, collections.Counter(), collections.Counter()
    idle = 0.0
    for t, d, i in pts:
        dt = t - last
        if dt > 0:
            level_time[len(live)] += dt
            if len(live) == 1:
                solo[ker[next(iter(live))][2][:70]] += dt
            if len(live) == 0:
                idle += dt
        last = t
        (live.add if d > 0 else live.discard)(i)
    busy = sum(e - s for s, e, _ in ker) / 1e3
    lines = [f"{which}: one graph replay; {len(ker)} kernels, span {span:.2f} ms under the profiler (plain replay {plain_ms:.2f} ms), "
             f"sum of kernel durations {busy:.2f} ms -> average {busy / span:.2f} kernels in flight",
             "time at each concurrency level (kernels in flight : ms : share of span):"]
    for lv in sorted(level_time):
        lines.append(f"  {lv:2d} : {level_time[lv] / 1e3:7.3f} : {100 * level_time[lv] / 1e3 / span:5.1f} %")
    lines.append("kernels that ran ALONE (the serial part), by total solo time:")
    for n, t in solo.most_common(14):
        lines.append(f"  {t / 1e3:7.3f} ms  {n}")
    byname = collections.Counter()
    for s, e, n in ker:
        byname[n[:70]] += e - s
    lines.append("kernel time by name (all streams):")
    for n, t in byname.most_common(10):
        lines.append(f"  {t / 1e3:7.3f} ms  {n}")
    nb = 40
    lines.append(f"timeline in {nb} buckets of {span / nb:.2f} ms: average kernels in flight | the two kernels with most time in the bucket")
    for b in range(nb):
        lo, hi = t0 + (t1 - t0) * b / nb, t0 + (t1 - t0) * (b + 1) / nb
        acc = collections.Counter()
        for s_, e_, n in ker:
            ov = min(e_, hi) - max(s_, lo)
            if ov > 0:
                acc[n.split("(")[0].replace("void ", "").replace("xva::", "").replace("anonymous namespace)::", "")[:44] + ("<" + n.split("<")[1].split(">")[0] + ">" if "gemm_tc_kernel<" in n else "")] += ov
        tot = sum(acc.values())
        top = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in acc.most_common(2)) if tot else ""
        lines.append(f"  {b:2d} {tot / (hi - lo):5.2f} | {top}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
