"""HiFi-GAN v1 training step (BASELINE.json configs[2]: G + MPD + MSD, batch 16, 8192-sample segments) through the
B200-native engine: audio-samples/s with the batch resident in HBM. Secondary metric (bench.py's JSON line is the
FastPitch configs[1] workload); prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from oracle import hifigan as ohg           # synthetic batch generator only
from xva_trainer_b200 import capi, hifigan as hg


class H(dict):
    __getattr__ = dict.__getitem__


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    h = H(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, learning_rate=2e-4, adam_b1=0.8,
          adam_b2=0.99, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
          fmax_for_loss=None)
    G = hg.Generator(h, device="cuda:0"); G.train()
    mpd = hg.MultiPeriodDiscriminator(device="cuda:0"); mpd.train()
    msd = hg.MultiScaleDiscriminator(device="cuda:0"); msd.train()
    step = hg.HiFiGANStep(G, mpd, msd, h)
    x, y, y_mel = (t.cuda() for t in ohg.synthetic_batch(B, 32, seed=1))
    use_graph = os.environ.get("XVA_NO_GRAPH") is None
    if use_graph:
        from xva_trainer_b200 import graph
        step.optim_g.lr_on_device = step.optim_d.lr_on_device = True
        gs = graph.GraphedStep(lambda a, b, c: step.step(a, b, c), [x, y, y_mel], warmup=3)
        run = lambda: gs()
    else:
        run = lambda: step.step(x, y, y_mel)
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        out = run()
    e1.record()
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"metric": "audio-samples/s (HiFi-GAN v1 G+MPD+MSD train step)", "value": B * 8192 / (ms * 1e-3),
                      "unit": "samples/s", "ms_per_step": ms, "host_enqueue_ms_per_step": host_s * 1e3 / steps, "batch": B,
                      "segment": 8192, "gpu_launches_per_step": capi.launch_count() / steps, "cuda_graph": use_graph,
                      "losses": {k: float(v) for k, v in out.items()}}))


if __name__ == "__main__":
    main()
