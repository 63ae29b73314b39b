"""Data parallelism of the FastPitch step on the CPU: the checks of tests/test_ddp_nccl_gpu.py -- which needs two
GPUs and is therefore skipped on a one-GPU box -- with two gloo ranks, each running the product package's FastPitch through
the emulated C ABI (tests/cabi_emu.py) on its own RAGGED shard of one batch:

  * with FastPitchLoss.set_distributed (global mask sums) + GradSync(mean=False) (SUM all-reduce issued slice by slice during
    backward) every rank ends up with the loss and the gradient arena of ONE process running the whole batch -- the
    reference's semantics, whose criterion runs on the outputs nn.DataParallel gathered (xva_train.py:790) -- while the mean
    of per-rank ratios (plain DDP averaging) is measurably different on the same batch;
  * after 3 optimizer steps with dropout on and per-rank dropout streams the parameter arenas of the two ranks are
    bit-identical without any broadcast.
What differs from the hardware test: gloo instead of NCCL, exact fp32 arithmetic instead of tf32 tensor cores. CPU only."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP_PATCHES = [('if self.device_.type != "cuda":', "if False:"),
              ('self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")',
               'self.device_ = torch.device("cpu")')]


def _shard(x, lo, hi):
    return [t[lo:hi].contiguous() if torch.is_tensor(t) else (t[lo:hi] if isinstance(t, list) else t) for t in x]


def _worker(rank, world, port, stage, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cabi_emu
    from oracle import fastpitch as ofp        # seeded weights / batch only
    from xva_trainer_b200 import parallel

    B, Tt, Tm = 4, 14, 44
    x, _ = ofp.synthetic_batch(B, Tt, Tm, seed=11, ragged=True)
    hl = (Tm, int(x[3].max()))                 # every run pads to the GLOBAL lengths, as nn.DataParallel's replicas do
    sd = ofp.make_state(1234)
    per = B // world
    targets = lambda xs: [xs[2], xs[1], xs[3], xs[9]]
    res = {}
    with cabi_emu.installed():
        fp = cabi_emu.load_module("fastpitch", FP_PATCHES)

        def make(seed_offset=0):
            m = fp.FastPitch(device="cpu")
            m.load_state_dict({k: v.clone() for k, v in sd.items()})
            m.training_stage = stage
            m.train()
            m.p_drop = 0.0
            m.seed = 1234 + seed_offset
            c = fp.FastPitchLoss()
            c.training_stage = stage
            return m, c

        xs = _shard(x, rank * per, (rank + 1) * per)
        # (1) global normalisation: the criterion all-reduces {sum, count}; gradients SUM-reduced during backward
        m, c = make()
        c.set_distributed(world)
        sync = parallel.GradSync(m, world, mean=False, min_bucket_elems=1 << 18)
        loss, _ = c(m(xs, host_lens=hl), targets(xs))
        m.zero_grad()
        m.backward(c, 1.0, grad_sync=sync)
        sync.finish()
        g_global = m.arena.g.clone()
        res["buckets"] = sync.buckets_sent
        # (2) mean of per-rank ratios (what plain DDP would do)
        m2, c2 = make()
        sync2 = parallel.GradSync(m2, world, mean=True)
        loss2, _ = c2(m2(xs, host_lens=hl), targets(xs))
        m2.zero_grad()
        m2.backward(c2, 1.0, grad_sync=sync2)
        sync2.finish()
        g_mean = m2.arena.g.clone()
        # (3) one process, whole batch
        m3, c3 = make()
        loss3, _ = c3(m3(x, host_lens=hl), targets(x))
        m3.zero_grad()
        m3.backward(c3, 1.0)
        g_one = m3.arena.g.clone()
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
        res.update(loss_global=float(loss), loss_one=float(loss3), loss_local=float(loss2), g_global_vs_one=rel(g_global, g_one),
                   g_mean_vs_one=rel(g_mean, g_one))
        # (4) three optimizer steps, dropout on with per-rank streams: replicas stay bit-identical
        m4, c4 = make(seed_offset=rank)
        m4.p_drop = 0.1
        c4.set_distributed(world)
        opt = fp.Lamb(m4, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        sync4 = parallel.GradSync(m4, world, mean=False)
        for i in range(3):
            fp.adjust_learning_rate(50000 + i, opt, 0.1, 1000)
            m4.zero_grad()
            c4(m4(xs, host_lens=hl), targets(xs))
            m4.backward(c4, 1.0, grad_sync=sync4)
            sync4.finish()
            opt.step()
            m4.step_dropout()
        p0 = m4.arena.p.clone()
        dist.broadcast(p0, src=0)
        res["replicas_identical"] = bool(torch.equal(p0, m4.arena.p))
        res["moved"] = rel(m4.arena.p, m.arena.p)
        res["finite"] = bool(torch.isfinite(m4.arena.p).all())
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process_on_the_whole_batch():
    stage = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, stage, q)) for r in range(2)]
    half = str(max(1, (os.cpu_count() or 2) // 2))                  # two ranks share the host: no BLAS oversubscription
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    os.environ.update({k: half for k in saved})
    try:
        for p in procs:
            p.start()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    got = dict(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        res = got[r]
        assert abs(res["loss_global"] - res["loss_one"]) <= 1e-6 * abs(res["loss_one"]), res
        assert res["g_global_vs_one"] < 2e-5, res          # summation order only
        assert res["g_mean_vs_one"] > 1e-3, res            # ragged shards: plain DDP averaging is measurably different
        assert res["replicas_identical"] and res["finite"] and res["moved"] > 1e-4, res
        assert res["buckets"] >= 2
    assert got[0]["loss_global"] == got[1]["loss_global"]
    assert got[0]["loss_local"] != got[1]["loss_local"]


# ------------------------------------------------------------------------------------- next tier: xVAPitch text encoder, pitch predictor
def _te_worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cabi_emu
    from textenc_util import TE_PATCHES, fill_pitch, pitch_ref_spec, seeded_state
    from xva_trainer_b200 import parallel

    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    layers, B, T, lens = 2, 4, 13, [13, 8, 11, 5]
    gen = torch.Generator().manual_seed(9)
    sd = seeded_state(layers=layers)
    tokens = torch.randint(1, 50, (B, T), generator=gen)
    lang = torch.randn(B, 12, generator=gen)
    rx = torch.randn(B, T, 204, generator=gen)
    psd = fill_pitch(pitch_ref_spec(layers=2), gen)
    px = torch.randn(B, T, 196, generator=gen)
    spk = torch.nn.functional.normalize(torch.randn(B, 512, 1, generator=gen), dim=1)
    pr = torch.randn(B, 1, T, generator=gen)
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    li = torch.tensor(lens, dtype=torch.int32)
    res = {}
    with cabi_emu.installed():
        te = cabi_emu.load_module("textenc", TE_PATCHES)
        hg = cabi_emu.load_module("hifigan", [('if dev.type != "cuda":', "if False:")])

        def encoder(p=0.0, seed=1234):
            m = te.TextEncoder(50, 192, 192, 768, 2, layers, 3, p, language_emb_dim=12, device="cpu", seed=seed)
            m.load_state_dict(sd)
            m.train()
            m.zero_grad()
            return m

        def enc_step(m, a, b, sync=None):
            m.forward_cl(tokens[a:b].contiguous(), li[a:b].contiguous(), lang[a:b].contiguous())
            m.backward_cl(rx[a:b].contiguous())
            if sync is not None:                            # backward finishes the arena top-down: last layer first
                tops = sorted({k.split(".")[0] for k in m.arena.offset}, reverse=True)
                for t in tops:
                    sync.ready([t])
                sync.finish()

        # (1) text encoder: SUM of the shard gradients == one process on the whole batch (the loss is a plain sum over items)
        m = encoder()
        sync = parallel.GradSync(m, world, mean=False, min_bucket_elems=1 << 16)
        enc_step(m, lo, hi, sync)
        one = encoder()
        enc_step(one, 0, B)
        res["te_vs_one"] = rel(m.flat.grad, one.flat.grad)
        res["te_buckets"], res["te_elems"], res["te_arena"] = sync.buckets_sent, sync.elems_sent, int(m.flat.numel())
        local = encoder()
        enc_step(local, lo, hi)
        res["te_local_vs_one"] = rel(local.flat.grad, one.flat.grad)
        # (2) pitch predictor (708 channels), mean=True: every rank back-propagates 1/world of its shard's loss
        def predictor():
            q = te.RelativePositioningPitchEnergyEncoder(1, 196, 768, 2, 2, 3, 0.0, conditioning_emb_dim=512, device="cpu")
            q.load_state_dict(psd)
            q.train()
            q.zero_grad()
            return q

        q = predictor()
        qs = parallel.GradSync(q, world, mean=True)
        q(px[lo:hi].contiguous(), lens[lo:hi], speaker_emb=spk[lo:hi].contiguous())
        q.backward(pr[lo:hi].contiguous() * qs.loss_scale)
        qs.ready(sorted({k.split(".")[0] for k in q.arena.offset}), flush=True)
        qs.finish()
        q1 = predictor()
        q1(px, lens, speaker_emb=spk)
        q1.backward(pr / world)
        res["pp_vs_one"] = rel(q.flat.grad, q1.flat.grad)
        # (3) two AdamW steps, dropout on with per-rank streams: replicas stay bit-identical without a broadcast
        m4 = encoder(p=0.1, seed=1234 + rank)
        opt = hg.AdamW([m4.flat], lr=1.75e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01)
        s4 = parallel.GradSync(m4, world, mean=False)
        for _ in range(2):
            m4.zero_grad()
            enc_step(m4, lo, hi, s4)
            opt.step()
            m4.step_dropout()
        p0 = m4.flat.data.clone()
        dist.broadcast(p0, src=0)
        res["replicas_identical"] = bool(torch.equal(p0, m4.flat.data))
        res["moved"] = rel(m4.flat.data, m.flat.data)
        res["finite"] = bool(torch.isfinite(m4.flat.data).all())
    out_q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_text_encoder_and_pitch_predictor_on_two_gloo_ranks():
    """The next tier's modules under the same exchange (textenc._RelTransformer.arena -> parallel.GradSync): two ranks on ragged
    shards of one batch end up with the gradient arena of one process on the whole batch, the exchange covers the whole arena
    in more than one bucket, and after two AdamW steps with per-rank dropout streams the replicas are bit-identical."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30800 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_te_worker, args=(r, 2, port, q)) for r in range(2)]
    half = str(max(1, (os.cpu_count() or 2) // 2))
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    os.environ.update({k: half for k in saved})
    try:
        for p in procs:
            p.start()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        res = got[r]
        assert res["te_vs_one"] < 2e-5 and res["pp_vs_one"] < 2e-5, res          # summation order only
        assert res["te_local_vs_one"] > 1e-2, res                                # a shard alone is not the batch
        assert res["te_buckets"] >= 2 and res["te_elems"] >= 0.99 * res["te_arena"], res
        assert res["replicas_identical"] and res["finite"] and res["moved"] > 1e-5, res
    print("rank 0:", got[0])
