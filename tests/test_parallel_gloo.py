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
