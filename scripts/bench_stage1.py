"""FastPitch training stage 1 (the aligner) at BASELINE.json's FastPitch configuration
(batch 32 x 880 frames x 160 tokens) through the B200-native engine: forward (ConvAttention, MAS), AttentionCTCLoss +
AttentionBinarizationLoss, backward, clip + LAMB. Prints one JSON line: mel-frames/s of the step with the batch resident
in HBM, the device time of each csrc/align.cu kernel (CUDA events on the launching stream, one instrumented step after
the timed region) with its algorithmic HBM bytes, and the CPU oracle's stage-1 step on a bounded sample next to it.

    python scripts/bench_stage1.py [steps] [--no-cpu]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from oracle import fastpitch as ofp          # synthetic batch generator and the CPU baseline only
from xva_trainer_b200 import capi, fastpitch as fp, ops

B, TT, TM, C = 32, 160, 880, 80


def algorithmic_bytes():
    """fp32 bytes each operator must move at least once: its inputs read once, its outputs written once. Workspaces and
    intermediates an implementation chooses to keep in HBM (the fp64 alpha / beta tables of the CTC kernels: 289 MB
    written and read back; the raw-score gradient dD between the two backward kernels: 18 MB each way) are NOT counted,
    so achieved / peak says how far each operator is from the one-pass HBM bound."""
    s = B * TM * TT * 4            # one score-sized tensor, 18.0 MB
    q, k = B * TM * C * 4, B * TT * C * 4
    return {"xva_attn_score_fwd": q + k + s + 2 * s,                 # q, k, prior -> logprob, soft
            "xva_mas_width1": s + s + B * TT * 4,                    # log-probabilities -> hard, durations
            "xva_mas_log": 2 * s,
            "xva_attn_ctc": 2 * s + B * 8,                           # logprob -> gradient, cost
            "xva_attn_bin_loss": 2 * s,
            "xva_attn_grad_combine": 3 * s + s,
            "xva_attn_score_bwd": 3 * s + q + k + q + k}             # g, logprob, prior, q, k -> dq, dk


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
    x_cpu, y_cpu = ofp.synthetic_batch(B, TT, TM, seed=1234, ragged=False, prior=True)
    frames = int(x_cpu[3].sum())
    dev = torch.device("cuda:0")
    x = [t.to(dev) if torch.is_tensor(t) else t for t in x_cpu]
    y = [x[2], x[1], x[3], x[9]]
    model = fp.FastPitch(device=dev, seed=1234)
    model.training_stage = 1
    model.train()
    crit = fp.FastPitchLoss()
    crit.training_stage = 1
    kl = fp.AttentionBinarizationLoss()
    opt = fp.Lamb(model, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)

    def step():
        model.zero_grad()
        out = model(x)
        loss, meta = crit(out, y)
        klv = kl(out[9], out[8])
        model.backward(crit, 1.0, kl=(kl, 0.5))
        opt.step()
        return loss + 0.5 * klv

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = capi.launch_count() / steps

    # instrumented step: events around every C-ABI call
    rec = []
    orig = capi.call

    def timed(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(name, *a)
        e.record()
        rec.append((name, s, e))

    capi.call = timed
    torch.cuda._sleep(int(0.05 * 1.9e9))
    step()
    torch.cuda.synchronize()
    capi.call = orig
    agg = {}
    for name, s, e in rec:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += s.elapsed_time(e)
    peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    hbm_peak = peaks.get("hbm_gbs") or 6547.5          # MEASURED_PEAKS.json (driver-written copy bandwidth), else the recipe's fallback
    ab = algorithmic_bytes()
    kernels = {}
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        k = {"launches": n, "ms": round(t, 4)}
        if name in ab:
            k["algorithmic_MB"] = round(ab[name] / 1e6, 1)
            k["achieved_GBps"] = round(ab[name] / (t * 1e-3) / 1e9, 1)
            k["frac_of_hbm_peak"] = round(ab[name] / (t * 1e-3) / 1e9 / hbm_peak, 3)
        kernels[name] = k
    line = {"metric": "mel-frames/s (FastPitch 1.1 stage-1 aligner step)", "value": frames / (ms * 1e-3), "unit": "frames/s",
            "ms_per_step": ms, "steps": steps, "gpu_launches_per_step": launches, "loss": float(loss),
            "config": {"workload": f"FastPitch1.1 stage 1 (aligner), batch={B}, {TM} frames x {TT} tokens, synthetic, "
                                   "beta-binomial prior, kl_weight 0.5", "launch": "eager"},
            "hbm_peak_GBps": hbm_peak, "kernels": kernels}
    if "--no-cpu" not in sys.argv:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        xb, yb = ofp.synthetic_batch(4, TT, TM, seed=1234, prior=True)
        sd = ofp.make_state(1234, perturb=False)
        times = []
        for i in range(3):
            t0 = time.perf_counter()
            ofp.train_step(sd, xb, yb, 1, ofp.noam_lr(50000 + i), {}, drop=0.0, training=True, kl_weight=0.5)
            times.append(time.perf_counter() - t0)
        per = sum(times[1:]) / 2
        line["cpu_baseline"] = {"value": int(xb[3].sum()) / per, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"oracle stage-1 train_step, batch 4 x {TM} frames of the batch-32 workload, 1 warm-up + 2 "
                                          f"timed steps ({per:.2f} s/step)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
