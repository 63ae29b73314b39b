"""What do the idle halves of fractional waves cost, and which scheduling change gets them back?  A discrete-event model
of 148 SMs running the persistent tap-GEMM launches of one FastPitch decoder-layer backward (the shapes and per-tile
times of profiles/r01_s5_fastpitch_gemm_table.txt), one CTA per SM (the kernel's ~200 KiB of shared memory):

  sequential          one stream, static round-robin tiles (what runs today)
  two streams static  input-gradient kernels on one stream, weight-gradient kernels on a second (XVA_BWD_STREAMS=1):
                      a CTA of the next kernel starts as soon as an SM is free, tiles stay statically assigned
  two streams dynamic the same with an atomic tile counter per launch (work-conserving)
  stream-K            one stream, every launch split evenly over the SMs along k (the bound: no idle SM, no overlap needed)

CPU only; prints the makespan of the sequence in units of one full-tile time of the first launch."""
import heapq

SMS = 148


def run(kernels, streams, dynamic):
    """kernels: list of (stream id, n_tiles, tile_time, after) in launch order; a kernel may start when every earlier kernel
    of its stream has been fully DISPATCHED (all its CTAs placed) -- CUDA's in-stream ordering is completion, so use the
    stricter rule: an in-stream successor starts after its predecessor has finished. `after`: index of a kernel in the
    other stream that must have finished first (data dependence) or None."""
    free = [(0.0, sm) for sm in range(SMS)]          # (time the SM becomes free, sm)
    heapq.heapify(free)
    done = [None] * len(kernels)
    last_in_stream = {}
    order = sorted(range(len(kernels)), key=lambda i: i)
    pending = list(order)
    t_end = 0.0
    while pending:
        # pick the launchable kernel with the earliest ready time
        best, best_ready = None, None
        for i in pending:
            s, n, d, after = kernels[i]
            prev = last_in_stream.get(s)
            deps = [p for p in (prev, after) if p is not None]
            if any(done[p] is None for p in deps):
                continue
            ready = max([done[p] for p in deps], default=0.0)
            if best is None or ready < best_ready:
                best, best_ready = i, ready
        i = best
        s, n, d, after = kernels[i]
        pending.remove(i)
        ctas = min(n, SMS)
        finish = 0.0
        if dynamic:
            remaining = n
            slots = []
            for _ in range(ctas):
                t, sm = heapq.heappop(free)
                slots.append((max(t, best_ready), sm))
            heapq.heapify(slots)
            while remaining:
                t, sm = heapq.heappop(slots)
                heapq.heappush(slots, (t + d, sm))
                remaining -= 1
            for t, sm in slots:
                heapq.heappush(free, (t, sm))
                finish = max(finish, t)
        else:
            for c in range(ctas):
                t, sm = heapq.heappop(free)
                start = max(t, best_ready)
                my_tiles = len(range(c, n, ctas))
                end = start + my_tiles * d
                heapq.heappush(free, (end, sm))
                finish = max(finish, end)
        done[i] = finish
        last_in_stream[s] = i
        t_end = max(t_end, finish)
    return t_end


def layer_backward(two_streams):
    """One decoder FFT block backward at B = 32 x 880: (stream, tiles, tile time) with tile time in units of the N = 384
    conv tile (205 us / 2 waves). dgrad chain on stream 0; weight gradients on stream 1 when two_streams."""
    w = 1 if two_streams else 0
    # name, tiles, relative tile time (measured us per launch / waves / reference tile)
    return [
        (w, 270, 0.55, None),     # wgrad W2: 3 taps x 3 x 6 n-tiles x split 5 (182 us, 1.82 waves)
        (0, 224 * 6, 0.083, None),  # dgrad through W2 -> 1536 columns as 6 n tiles of 256 (186 us)
        (w, 216, 0.72, 1),        # wgrad W1 (155 us, 1.46 waves) needs dh from the dgrad above
        (0, 224, 0.86, 1),        # dgrad through W1 -> 384 columns, one 384-wide tile per row tile (176 us, 1.51 waves)
    ]


def main():
    seq = layer_backward(False)
    base = run(seq, 1, False)
    print(f"sequential, static       : {base:6.2f}")
    print(f"two streams, static      : {run(layer_backward(True), 2, False):6.2f}")
    print(f"two streams, dynamic     : {run(layer_backward(True), 2, True):6.2f}")
    work = sum(n * d for _, n, d, _ in seq) / SMS
    print(f"stream-K bound (no idle) : {work:6.2f}")
    print(f"idle fraction today      : {1 - work / base:6.1%}")


if __name__ == "__main__":
    main()
