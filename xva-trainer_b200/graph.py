"""CUDA-graph replay of a whole training step.

A step of this engine is a fixed sequence of ~330 (FastPitch) to ~5 000 (HiFi-GAN) kernel launches with static shapes;
issued one by one from Python it is bound by host launch cost long before the GPU is. ``GraphedStep`` warms the step up,
captures it once (every launch goes through the C ABI on torch's current stream, so capture sees them; TMA descriptors
travel as kernel parameters and the graph's private memory pool pins every address) and afterwards replays it with one
``cudaGraphLaunch``. Per-step host inputs are copied into the static tensors the captured step reads; quantities that
change from step to step live on the device (dropout counter, AdamW step count, learning rate)."""
import torch


class GraphedStep:
    def __init__(self, fn, static_inputs, warmup=3, pool=None):
        """fn(*static_inputs) -> pytree of tensors; static_inputs: the tensors whose STORAGE the captured step reads.
        ``warmup`` eager executions precede the capture (0 when the caller has already run the step: a training step
        has side effects). ``pool``: the memory pool of another GraphedStep whose replays never overlap this one's
        (graphs of different batch shapes of one trainer), so their activations share memory."""
        self.fn, self.static_inputs = fn, list(static_inputs)
        if warmup > 0:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fn(*self.static_inputs)
            cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.outputs = fn(*self.static_inputs)
        self.pool = self.graph.pool()

    def load(self, *new_inputs):
        """Copy fresh values into the static input tensors (async on the current stream)."""
        for dst, src in zip(self.static_inputs, new_inputs):
            if dst is not None and src is not None and dst is not src:
                dst.copy_(src, non_blocking=True)

    def __call__(self, *new_inputs):
        if new_inputs:
            self.load(*new_inputs)
        self.graph.replay()
        return self.outputs
