"""Data-parallel gradient exchange (xva-trainer_b200/parallel.py) on CPU: two gloo ranks, a stand-in gradient arena with
the FastPitch layout, ready() calls in the order backward() issues them. Checks that every touched slice ends up as the
sum over ranks, that buckets are merged, and that nothing outside the announced slices is modified."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeArena:
    def __init__(self, rank):
        names = (["encoder.word_emb.weight"] + [f"encoder.layers.{i}.w" for i in range(3)] + ["duration_predictor.w"]
                 + [f"decoder.layers.{i}.w" for i in range(3)] + ["pitch_predictor.w", "proj.weight", "attention.w"])
        self.offset, self.pshape = {}, {}
        off = 0
        for n in names:
            self.offset[n], self.pshape[n] = off, (1000,)
            off += 1024  # alignment gaps, like the real arena
        self.g = torch.full((off,), float(rank + 1))


def _worker_stage1(rank, world, port, q):
    """Stage 1 (the aligner): backward() finishes exactly two slices at opposite ends of the arena -- the attention
    projection stacks and the token embedding -- and announces them separately (fastpitch.py::_backward_stage1)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xva_trainer_b200.parallel import GradSync

    arena = _FakeArena(rank)
    sync = GradSync(arena, world)
    sync.ready(["attention"], flush=True)
    sync.ready(["encoder.word_emb"], flush=True)
    sync.finish()
    total, own = float(sum(r + 1 for r in range(world))), float(rank + 1)
    ok = True
    for name, off in arena.offset.items():
        touched = name.startswith("attention") or name.startswith("encoder.word_emb")
        ok &= bool(torch.all(arena.g[off:off + 1000] == (total if touched else own)))
    q.put((rank, ok, sync.buckets_sent, sync.elems_sent))
    dist.destroy_process_group()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xva_trainer_b200.parallel import GradSync

    arena = _FakeArena(rank)
    sync = GradSync(arena, world, min_bucket_elems=1500)
    assert sync.loss_scale == 1.0 / world
    for i in (2, 1, 0):
        sync.ready([f"decoder.layers.{i}"])
    sync.ready(["pitch_predictor", "proj"], flush=True)
    for i in (2, 1):
        sync.ready([f"encoder.layers.{i}"])
    sync.ready(["encoder.layers.0", "encoder.word_emb"], flush=True)
    sync.finish()
    total = float(sum(r + 1 for r in range(world)))
    own = float(rank + 1)
    ok = True
    for name, off in arena.offset.items():
        got = arena.g[off:off + 1000]
        touched = not (name.startswith("duration_predictor") or name.startswith("attention"))
        ok &= bool(torch.all(got == (total if touched else own)))
    q.put((rank, ok, sync.buckets_sent))
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, buckets in res:
        assert ok, f"rank {rank}: all-reduced arena is wrong"
        assert 2 <= buckets <= 6, buckets  # adjacent layer slices were merged into fewer, larger messages


def test_gradsync_stage1_slices_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_stage1, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, buckets, elems in res:
        assert ok, f"rank {rank}: all-reduced arena is wrong"
        assert buckets == 2 and elems == 2000, (buckets, elems)   # only the two slices travel, not the arena between them


def test_shard_batches_is_strided_and_drops_the_remainder():
    sys.path.insert(0, ROOT)
    from xva_trainer_b200.parallel import shard_batches

    items = list(range(11))
    shards = [shard_batches(items, r, 4) for r in range(4)]
    assert shards == [[0, 4], [1, 5], [2, 6], [3, 7]]                 # 11 // 4 = 2 per rank, 3 items dropped
    assert shard_batches(items, 0, 1) == items


def _worker_mask_sums(rank, world, port, q):
    """The criterion's {sum, count} pairs are all-reduced before the ratios: every rank then reports the loss of the
    global batch, sum_global(err * mask) / sum_global(mask), not its local ratio."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xva_trainer_b200.fastpitch import AttentionBinarizationLoss, FastPitchLoss

    crit = FastPitchLoss().set_distributed(world)
    acc = torch.tensor([[10.0 * (rank + 1), 100.0 + 50 * rank], [0.0, 0.0], [3.0 + rank, 7.0 - rank], [1.0, 2.0]], dtype=torch.float64)
    crit._all_reduce(acc)
    kl = AttentionBinarizationLoss().set_distributed(world)
    q.put((rank, acc.tolist(), kl.world, FastPitchLoss().world))
    dist.destroy_process_group()


def test_loss_mask_sums_are_global_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_worker_mask_sums, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[30.0, 250.0], [0.0, 0.0], [7.0, 13.0], [2.0, 4.0]]
    for rank, acc, klw, default_world in out:
        assert acc == want and klw == 2 and default_world == 1
    # the global ratio differs from the mean of the per-rank ratios whenever the counts differ
    assert abs(30.0 / 250.0 - 0.5 * (10.0 / 100.0 + 20.0 / 150.0)) > 1e-3


def _worker_pad(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xva_trainer_b200.parallel import pad_batches_to_global

    def batch(B, Tt, Tm, prior):
        x = [torch.ones(B, Tt, dtype=torch.long), torch.full((B,), Tt), torch.ones(B, 80, Tm), torch.full((B,), Tm),
             torch.ones(B, 1, Tm), torch.ones(B, Tm), None, torch.ones(B, Tm, Tt) if prior else None, torch.ones(B, Tt),
             torch.full((B,), float(Tt)), torch.full((B,), float(Tm)), ["u"] * B]
        return (x, [x[2], x[1], x[3], x[9]], B * Tm)

    mine = [batch(2, 10 + rank, 50 - 3 * rank, True), batch(2, 12, 40, False)]
    out = pad_batches_to_global(mine, world)
    shapes = [(tuple(b[0][0].shape), tuple(b[0][2].shape), tuple(b[0][4].shape), tuple(b[0][5].shape),
               None if b[0][7] is None else tuple(b[0][7].shape), tuple(b[0][8].shape), float(b[0][9][0]), float(b[0][10][0]),
               b[1][0] is b[0][2], int(b[0][1][0]), int(b[0][3][0]), float(b[0][2].sum()), float(b[0][0].sum())) for b in out]
    q.put((rank, shapes))
    dist.destroy_process_group()


def test_pad_batches_to_global_gloo():
    """Every rank's i-th batch ends up with the maximum text / mel length over ranks; lengths, sums (zero padding) and the
    target list follow; a batch that is already at the maximum is returned untouched."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 11
    procs = [ctx.Process(target=_worker_pad, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        first, second = out[rank]
        Tt, Tm = 10 + rank, 50 - 3 * rank
        assert first[:6] == ((2, 11), (2, 80, 50), (2, 1, 50), (2, 50), (2, 50, 11), (2, 11))
        assert first[6:9] == (11.0, 50.0, True)
        assert first[9:11] == (Tt, Tm)                                   # the true lengths are not touched
        assert first[11] == 2 * 80 * Tm and first[12] == 2 * Tt            # padding is zeros
        assert second[:2] == ((2, 12), (2, 80, 40)) and second[4] is None
