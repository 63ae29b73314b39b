"""Data parallelism on real hardware: two NCCL ranks, one per GPU, each on its own RAGGED shard of one batch.

  * loss and gradients: with FastPitchLoss.set_distributed (global mask sums) + GradSync(mean=False) (SUM all-reduce
    overlapped with backward) every rank ends up with the loss and the gradient arena of ONE process running the whole
    batch -- the reference's semantics, whose criterion runs on the outputs nn.DataParallel gathered (xva_train.py:790).
    The plain mean of per-rank ratios (GradSync(mean=True) alone) is shown to differ on the same ragged batch.
  * replicas: after 5 optimizer steps with dropout on (per-rank streams) the parameter arenas of the two ranks are
    bit-identical without any broadcast.
Needs >= 2 GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl_gpu.py -m gpu`.
"""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(x, lo, hi):
    return [t[lo:hi].contiguous() if torch.is_tensor(t) else (t[lo:hi] if isinstance(t, list) else t) for t in x]


def _worker(rank, world, port, stage, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import __graft_entry__ as ge

    ge.build()
    from oracle import fastpitch as ofp        # seeded weights / batch only
    from xva_trainer_b200 import fastpitch as fp, parallel

    B, Tt, Tm = 4, 40, 150
    x, _ = ofp.synthetic_batch(B, Tt, Tm, seed=11, ragged=True)
    # every run pads to the GLOBAL lengths, as nn.DataParallel's replicas do (the shards below are slices of one padded
    # batch; the decoder length is handed in): an utterance's last frames depend on the padded length of its batch --
    # nothing masks between the two convolutions of PositionwiseConvFF -- so local padding would differ from the
    # whole-batch run by ~1e-5 in the loss (measured: profiles/r02_ddp_nccl_stage3.json, r02_batch_invariance.txt)
    hl = (Tm, int(x[3].max()))
    sd = ofp.make_state(1234)
    per = B // world
    to_dev = lambda xs: [t.to(dev) if torch.is_tensor(t) else t for t in xs]
    targets = lambda xs: [xs[2], xs[1], xs[3], xs[9]]

    def make(seed_offset=0):
        m = fp.FastPitch(device=dev)
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.training_stage = stage
        m.train()
        m.p_drop = 0.0
        m.seed = 1234 + seed_offset
        c = fp.FastPitchLoss()
        c.training_stage = stage
        return m, c

    res = {}
    xs = to_dev(_shard(x, rank * per, (rank + 1) * per))
    # ---- (1) global normalisation: criterion all-reduces {sum, count}; gradients SUM-reduced during backward
    m, c = make()
    c.set_distributed(world)
    sync = parallel.GradSync(m, world, mean=False, min_bucket_elems=1 << 18)
    loss, meta = c(m(xs, host_lens=hl), targets(xs))
    m.zero_grad()
    m.backward(c, 1.0, grad_sync=sync)
    sync.finish()
    torch.cuda.synchronize()
    g_global = m.arena.g.clone()
    res["buckets"] = sync.buckets_sent
    # ---- (2) mean of per-rank ratios (what plain DDP would do)
    m2, c2 = make()
    sync2 = parallel.GradSync(m2, world, mean=True)
    loss2, _ = c2(m2(xs, host_lens=hl), targets(xs))
    m2.zero_grad()
    m2.backward(c2, 1.0, grad_sync=sync2)
    sync2.finish()
    g_mean = m2.arena.g.clone()
    # ---- (3) one process, whole batch (rank 0 computes, everyone compares against its broadcast)
    m3, c3 = make()
    xf = to_dev(x)
    loss3, _ = c3(m3(xf, host_lens=hl), targets(xf))
    m3.zero_grad()
    m3.backward(c3, 1.0)
    torch.cuda.synchronize()
    g_one = m3.arena.g.clone()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    res.update(loss_global=float(loss), loss_one=float(loss3), loss_local=float(loss2), g_global_vs_one=rel(g_global, g_one),
               g_mean_vs_one=rel(g_mean, g_one))
    # ---- (4) five optimizer steps, dropout on with per-rank streams: replicas stay bit-identical
    m4, c4 = make(seed_offset=rank)
    m4.p_drop = 0.1
    c4.set_distributed(world)
    opt = fp.Lamb(m4, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    sync4 = parallel.GradSync(m4, world, mean=False)
    for i in range(5):
        fp.adjust_learning_rate(50000 + i, opt, 0.1, 1000)
        m4.zero_grad()
        l4, _ = c4(m4(xs, host_lens=hl), targets(xs))
        m4.backward(c4, 1.0, grad_sync=sync4)
        sync4.finish()
        opt.step()
        m4.step_dropout()
    torch.cuda.synchronize()
    p0 = m4.arena.p.clone()
    dist.broadcast(p0, src=0)
    res["replicas_identical"] = bool(torch.equal(p0, m4.arena.p))
    res["moved"] = rel(m4.arena.p, m.arena.p)
    res["finite"] = bool(torch.isfinite(m4.arena.p).all())
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("stage", [3, 2])
def test_two_ranks_equal_one_process_on_the_whole_batch(stage):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000) + stage
    procs = [ctx.Process(target=_worker, args=(r, 2, port, stage, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import json
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ddp_nccl_stage{stage}.json"), "w") as f:
        json.dump(got, f, indent=1)
    print("ddp results:", got)
    for r in (0, 1):
        res = got[r]
        # the global-normalised loss is the whole-batch loss on every rank (fp64 reductions: 1e-6 covers tf32-level
        # differences of T_out padding between the runs), the mean of per-rank ratios is not
        assert abs(res["loss_global"] - res["loss_one"]) <= 1e-6 * abs(res["loss_one"]), res
        assert res["g_global_vs_one"] < 2e-5, res          # order of fp32 atomic additions only
        assert res["g_mean_vs_one"] > 1e-3, res            # ragged shards: plain DDP averaging is measurably different
        assert res["replicas_identical"] and res["finite"] and res["moved"] > 1e-4, res
        assert res["buckets"] >= 2
    assert got[0]["loss_global"] == got[1]["loss_global"]
    assert got[0]["loss_local"] != got[1]["loss_local"]


def _vits_worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import __graft_entry__ as ge

    ge.build()
    from test_oracle_golden import _vits_hifi_only_fixture
    from xva_trainer_b200 import hifigan as hg, vits

    gold, specs, sds, linear, waveform, d_vectors = _vits_hifi_only_fixture()
    eps, u = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["u"])
    lens = [int(v) for v in gold["y_lengths"]]

    def make(world_):
        enc = vits.PosteriorEncoder(513, 192, 192, kernel_size=5, dilation_rate=1, num_layers=16, cond_channels=512, device=dev)
        dec = hg.HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                                  inference_padding=0, cond_channels=512, conv_pre_weight_norm=False,
                                  conv_post_weight_norm=False, conv_post_bias=False, device=dev)
        disc = hg.VitsDiscriminator(device=dev)
        for name, mod in (("enc", enc), ("dec", dec), ("disc", disc)):
            mod.load_state_dict(sds[name])
            mod.train()
        return vits.HifiOnlyStep(enc, dec, disc, world=world_)

    # two ranks, one utterance each (the second padded to the batch's 40 frames, as a DataParallel replica sees it)
    st = make(world)
    sl = slice(rank, rank + 1)
    losses = st.step(linear[sl], lens[rank:rank + 1], waveform[sl], d_vectors[sl], eps=eps[sl], u=u[sl])
    torch.cuda.synchronize()
    flat = lambda s_: torch.cat([p.detach().reshape(-1) for p in s_.optim_g.params + s_.optim_d.params]).clone()
    mine = flat(st)
    other = mine.clone()
    dist.broadcast(other, src=0)
    res = {"replicas_equal": bool(torch.equal(mine, other)), "loss_disc": float(losses["loss_disc"]), "loss": float(losses["loss"])}
    if rank == 0:
        one = make(1)
        l1 = one.step(linear, lens, waveform, d_vectors, eps=eps, u=u)
        torch.cuda.synchronize()
        ref = flat(one)
        res["update_err"] = float((mine - ref).double().norm() / (ref - torch.cat([t.reshape(-1) for sd in (sds["enc"], sds["dec"], sds["disc"])
                                                                                     for t in sd.values()]).to(dev)).double().norm())
        res["param_err"] = float((mine - ref).double().norm() / ref.double().norm())
        res["whole_batch"] = {k: float(l1[k]) for k in ("loss", "loss_disc")}
    ls = torch.tensor([res["loss"], res["loss_disc"]], device=dev, dtype=torch.float64)
    dist.all_reduce(ls)
    res["mean_losses"] = (ls / world).tolist()
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_vits_hifi_only_two_ranks_equal_one_process():
    """xVAPitch --hifi_only step on 2 NCCL ranks x 1 utterance vs 1 process x 2 utterances (the recorded fixture): the
    mean of the per-rank losses equals the whole-batch loss, both ranks hold identical parameters after the step without
    a broadcast, and they are the one-process parameters up to tf32 / summation-order noise on a sign-like first AdamW
    step (same bound as the oracle comparison of tests/test_vits_gpu.py)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 7
    procs = [ctx.Process(target=_vits_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0]["replicas_equal"] and got[1]["replicas_equal"]
    r0 = got[0]
    for i, k in enumerate(("loss", "loss_disc")):
        assert abs(r0["mean_losses"][i] - r0["whole_batch"][k]) < 2e-3 * abs(r0["whole_batch"][k]), (k, r0)
    assert r0["param_err"] < 5e-3, r0
    import json
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(got[0], open(os.path.join(ROOT, "gpurun_out", "ddp_nccl_vits.json"), "w"), indent=1)
