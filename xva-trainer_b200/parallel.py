"""Data-parallel gradient exchange: one process per GPU, the utterance batch sharded across ranks, and ONE collective
on the path -- an all-reduce (SUM) of the gradient arena before the optimizer step (the reference
itself only has single-process nn.DataParallel, xva_train.py:465-466).

The gradient arena is one flat fp32 buffer laid out in state_dict order, and backward() finishes it from the top of the
network down, so the exchange is bucketed by construction: as soon as a contiguous slice is final (a decoder layer, the
predictor / projection tail, an encoder layer) backward() calls ``ready(prefixes)`` and that slice is all-reduced in
place on NCCL's own stream while the remaining backward kernels keep the SMs busy. ``finish()`` joins the streams before
LAMB reads the arena. Every rank applies the same update to the same all-reduced gradients, so replicas stay
bit-identical without a broadcast.

Loss normalisation. ``mean=True``: loss gradients are pre-scaled by 1/world and SUM yields the mean of the
per-rank gradients -- exact for losses that are plain means over equal-sized shards. FastPitch's masked MSE terms are
ratios sum(err * mask) / sum(mask) and the reference evaluates them on the GATHERED outputs of all GPUs
(xva_train.py:790), which the mean of per-rank ratios only equals when every rank holds the same number of valid frames
and tokens. ``mean=False`` + ``FastPitchLoss.set_distributed(world)``: the criterion all-reduces its {sum, count} pairs
first, every rank back-propagates d(global loss)/d(its predictions), and the SUM of the gradients is exactly the gradient
of one GPU running the global batch.
"""
import math

import torch
import torch.distributed as dist


class GradSync:
    MAX_GAP = 256  # elements; the arena aligns tensors to 64 floats

    def __init__(self, model_or_arena, world, group=None, min_bucket_elems=1 << 20, mean=True):
        self.arena = getattr(model_or_arena, "arena", model_or_arena)
        self.world = int(world)
        self.group = group
        self.min_bucket = int(min_bucket_elems)
        self.mean = bool(mean)
        self.pending = []
        self._lo = None
        self._hi = None
        self.buckets_sent = 0
        self.elems_sent = 0

    @property
    def loss_scale(self):
        return 1.0 / self.world if self.mean else 1.0

    def _range(self, prefixes):
        A = self.arena
        lo, hi = None, None
        for k in A.offset:
            if any(k == p or k.startswith(p + ".") for p in prefixes):
                a = A.offset[k]
                b = a + int(math.prod(A.pshape[k]))
                lo = a if lo is None else min(lo, a)
                hi = b if hi is None else max(hi, b)
        if lo is None:
            raise KeyError(f"no arena entry under {prefixes}")
        return lo, hi

    def ready(self, prefixes, flush=False):
        """The gradients of every parameter under ``prefixes`` are final. Adjacent ready slices are merged until a
        bucket reaches min_bucket elements (or ``flush``), then all-reduced asynchronously."""
        lo, hi = self._range(prefixes)
        # slices separated only by the arena's alignment padding (never written, always zero) count as adjacent
        if self._lo is not None and lo <= self._hi + self.MAX_GAP and hi >= self._lo - self.MAX_GAP:
            self._lo, self._hi = min(lo, self._lo), max(hi, self._hi)
        else:
            self._send()
            self._lo, self._hi = lo, hi
        if flush or (self._hi - self._lo) >= self.min_bucket:
            self._send()

    def _send(self):
        if self._lo is None:
            return
        sl = self.arena.g[self._lo:self._hi]
        if self.world > 1:
            self.pending.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.buckets_sent += 1
        self.elems_sent += self._hi - self._lo
        self._lo = self._hi = None

    def finish(self):
        """Join: the current stream waits for every outstanding all-reduce."""
        self._send()
        for w in self.pending:
            w.wait()
        self.pending = []


def shard_batches(batches, rank, world):
    """Rank r takes items r, r + world, ... of the (already shuffled) list and every rank the same count (drop_last, as the
    reference's DataLoader at xva_train.py:452): ranks must issue the same number of collectives per epoch."""
    batches = list(batches)
    n = (len(batches) // world) * world
    return batches[rank:n:world]


def pad_batches_to_global(batches, world, group=None):
    """Pad the i-th batch of every rank to the same text and mel lengths (the maxima over ranks), in place of the
    per-rank padding a collate produces. Why it matters for parity: in the reference nothing masks between the two
    convolutions of PositionwiseConvFF (transformer.py:59-77), so the last valid frame of an utterance reads one frame of
    its own padding -- a frame that exists when the utterance is shorter than its batch and is a zero boundary when it is
    the longest. An utterance's output therefore depends (at the 1e-5 level in the loss) on the padded length of its
    batch. nn.DataParallel pads the GLOBAL batch before scattering it; ranks that pad locally would differ from that, and
    from one GPU running the global batch, by exactly this effect. ``batches``: list of (x, y, num_frames) in the layout
    of batch_to_gpu (data_function.py:706-741). One all-reduce(MAX) of 2 ints per batch, once, at trainer start."""
    import torch.nn.functional as F
    if world <= 1 or not batches:
        return batches
    dev = batches[0][0][0].device
    sizes = torch.tensor([[b[0][0].shape[1], b[0][2].shape[2]] for b in batches], device=dev, dtype=torch.int64)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
    out = []
    for (x, y, n), (Tt, Tm) in zip(batches, sizes.tolist()):
        dt, dm = Tt - x[0].shape[1], Tm - x[2].shape[2]
        if dt or dm:
            x = list(x)
            x[0] = F.pad(x[0], (0, dt))                                   # text [B, Tt]
            x[2] = F.pad(x[2], (0, dm))                                   # mel [B, 80, Tm]
            if x[4] is not None:
                x[4] = F.pad(x[4], (0, dm))                               # pitch [B, 1, Tm]
            if x[5] is not None:
                x[5] = F.pad(x[5], (0, dm))                               # energy [B, Tm]
            if x[7] is not None:
                x[7] = F.pad(x[7], (0, dt, 0, dm))                        # attn_prior [B, Tm, Tt]
            if x[8] is not None:
                x[8] = F.pad(x[8], (0, dt))                               # durations [B, Tt]
            x[9] = torch.full_like(x[9], float(Tt))                       # max_inp_lengths
            x[10] = torch.full_like(x[10], float(Tm))                     # max_mel_lengths
            y = [x[2], x[1], x[3], x[9]]
        out.append((x, y, n))
    return out
