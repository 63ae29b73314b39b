"""Trainer facades: the entry points the reference's server / UI drive, re-hosted on the B200 engine.

Mirrors (same names, constructor arguments, attributes, websocket strings and on-disk outputs):
    python/models_manager.py:8-164            ModelsManager.sync_init_model / models / models_bank
    python/fastpitch1_1/xva_train.py:57-176   handleTrainer            :185-1081  FastPitchTrainer
    python/hifigan/xva_train.py:50-125        handleTrainer (HiFi-GAN) :132-649   HiFiTrainer

Kept from the reference: the ``data`` dict (dataset_path, output_path, checkpoint, batch_size, epochs_per_checkpoint,
force_stage, ...), ``"Set stage to: N "`` / ``"Finished training HiFi-GAN\\n"`` websocket messages, ``training.log``
(rewritten whole, as the UI tails it), ``graphs.json`` ({stages: {"1".."5": {loss, loss_delta, target_delta}}}),
checkpoint names and keys (FastPitch_checkpoint_{epoch}_{iter}.pt = {epoch, iteration, avg_loss_per_epoch,
training_stage, state_dict, optimizer}; {voice}.pt fp16 state dict; {voice}.json; hifi/g_{steps:08d}, hifi/do_{steps:08d},
{voice}.hg.pt), two checkpoints retained, the per-stage early-stopping rule (relative loss delta averaged over a span,
patience 3; xva_train.py:920-972), noam learning rate, gradient accumulation to ~256 items.

Replaced: the recursive ``await self.iteration()`` + "recursion depth" catch (xva_train.py:908-909, 719-721) is a plain
loop; stage changes are a ``StageFinished`` exception (a RuntimeError, so handleTrainer's contract is unchanged) instead
of a bare ``raise``; ``LOCAL_RANK`` is parsed as int (Appendix C-5 of SURVEY.md).

Out of scope: the wav / text dataset loaders. Batches come from ``data["batch_source"]``
(any iterable of reference-layout batches) or, with ``dataset_path = "synthetic:<B>x<Tt>x<Tm>x<items>"``, from the seeded
synthetic generator of SURVEY.md section 8(d).
"""
import asyncio
import datetime
import json
import math
import os
import time
import traceback

import torch

from . import fastpitch as fp
from . import graph
from . import hifigan as hg
from . import ops
from . import parallel


class StageFinished(RuntimeError):
    """Raised by finish_epoch when a training stage (or the whole run) is over; handleTrainer inspects the trainer's
    JUST_FINISHED_STAGE / END_OF_TRAINING flags exactly as the reference does after its bare ``raise``."""


def _now():
    t = str(datetime.datetime.now().time())
    return t.split(".")[0]


def beta_binomial_prior_distribution(phoneme_count, mel_count, scaling_factor=1.0):
    """data_function.py:85-99 (scipy.stats.betabinom(P, a, b).pmf(0..P-1) per mel frame) from lgamma, [M, P] fp32."""
    i = torch.arange(1, mel_count + 1, dtype=torch.float64)[:, None]
    k = torch.arange(0, phoneme_count, dtype=torch.float64)[None, :]
    n = torch.tensor(float(phoneme_count), dtype=torch.float64)
    a, b = scaling_factor * i, scaling_factor * (mel_count + 1 - i)
    lg = torch.lgamma
    log_beta = lambda u, v: lg(u) + lg(v) - lg(u + v)
    return torch.exp(lg(n + 1) - lg(k + 1) - lg(n - k + 1) + log_beta(k + a, n - k + b) - log_beta(a, b)).float()


def get_target_delta(training_stage, num_data_lines):
    """FastPitchTrainer.get_target_delta, xva_train.py:588-672: the relative loss improvement per epoch below which a
    training stage counts as converged, by stage and dataset size (the thresholds and multipliers are the reference's;
    its first two stage-1 branches test the same bound, so 5e-5 is unreachable there as well). The parameter freezing the
    reference does in the same function is trainable_keys(stage) here."""
    n = num_data_lines
    if training_stage == 1:
        delta = 2e-5 if n > 4000 else 15e-5 if n > 2000 else 4e-4 if n > 500 else 0
        if n < 500:
            delta = 4e-4
        return delta
    if training_stage == 2:
        delta = 5e-5 if n > 4000 else 1e-4 if n > 2000 else 5e-4
        if n < 500:
            delta = 4e-3
        return delta * 1.5
    if training_stage == 3:
        delta = 5e-5 if n > 4000 else 1e-4 if n > 2000 else 6e-4
        if n < 500:
            delta = 2e-3 if n < 250 else 1e-3
        return delta * 2.5
    if training_stage == 4:
        delta = 35e-6 if n > 4000 else 1e-4 if n > 2000 else 25e-5
        if n < 500:
            delta = 15e-4 if n < 250 else 45e-5
        return delta * 1.5 * 2
    return 0


def _synthetic_fastpitch_batches(spec, device, seed=1234):
    """'synthetic:BxTtxTmxitems[:prior]' -> list of (x, y, num_frames) in the layout of batch_to_gpu
    (data_function.py:706-741). With ':prior' every batch carries the beta-binomial alignment prior (x[7]) and a run
    starts at training stage 1, like a new voice in the reference."""
    body = spec.split(":", 1)[1]
    with_prior = body.endswith(":prior")
    B, Tt, Tm, items = (int(v) for v in body.split(":")[0].split("x"))
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(max(1, items // B)):
        text = torch.randint(1, fp.N_SYMBOLS, (B, Tt), generator=g)
        durs = torch.ones(B, Tt)
        for b in range(B):
            durs[b] += torch.bincount(torch.randint(0, Tt, (Tm - Tt,), generator=g), minlength=Tt).float()
        mel = torch.randn(B, fp.N_MEL, Tm, generator=g)
        pitch = torch.randn(B, 1, Tm, generator=g) * (torch.rand(B, 1, Tm, generator=g) > 0.3)
        energy = torch.rand(B, Tm, generator=g) * 10
        lens = torch.full((B,), Tt, dtype=torch.long)
        mlens = torch.full((B,), Tm, dtype=torch.long)
        t = lambda v: v.to(device)
        prior = t(beta_binomial_prior_distribution(Tt, Tm).unsqueeze(0).repeat(B, 1, 1)) if with_prior else None
        x = [t(text), t(lens), t(mel), t(mlens), t(pitch), t(energy), None, prior, t(durs),
             t(torch.full((B,), float(Tt))), t(torch.full((B,), float(Tm))), ["synthetic"] * B]
        out.append((x, [x[2], x[1], x[3], x[9]], int(mlens.sum())))
    return out


class _TrainerBase:
    def __init__(self, logger, PROD, gpus, models_manager, websocket=None):
        self.logger = logger
        self.PROD = PROD
        self.models_manager = models_manager
        self.gpus = gpus
        # One process per GPU. The reference reaches multi-GPU as len(gpus) > 1 inside ONE process
        # (nn.DataParallel, xva_train.py:465-466); here the same ``gpus`` list is served by WORLD_SIZE processes launched
        # with torchrun, rank r driving gpus[LOCAL_RANK] -- LOCAL_RANK parsed as an int (the reference adds the string to
        # 1234 and dies, Appendix C-5).
        self.local_rank = int(os.getenv("LOCAL_RANK", 0))
        self.world = int(os.getenv("WORLD_SIZE", "1"))
        self.rank = int(os.getenv("RANK", "0"))
        gpu = gpus[self.local_rank] if (self.world > 1 and len(gpus) > self.local_rank) else (
            self.local_rank if self.world > 1 else gpus[0])
        self.device = torch.device(f"cuda:{gpu}")
        self.ckpt_path = None
        self.websocket = websocket
        self.training_log = []
        self.training_log_live_line = ""
        self.model = None
        self.isReady = True
        self.epoch = None
        self.running = False
        self.is_init = False
        self.logs_are_init = False
        self.dataset_id = self.dataset_input = self.dataset_output = None
        self.batch_size = self.force_stage = self.workers = None
        self.JUST_FINISHED_STAGE = False
        self.END_OF_TRAINING = False
        self.graphs_json = None

    # reference: xva_train.py:226-238 (rewrites the whole file: the UI polls it)
    def print_and_log(self, line=None, end="\n", flush=False, save_to_file=False):
        if line is not None:
            self.training_log.append(f"{_now()} | {line}")
        if save_to_file and self.rank == 0:
            with open(f"{save_to_file}/training.log", "w+") as f:
                f.write("\n".join(self.training_log + [self.training_log_live_line]))

    def load_state_dict(self, ckpt_path, sd):
        pass

    def set_device(self, device):
        pass

    def pause(self, websocket=None):
        self.logger.info("pause") if self.logger else None
        self.running = False

    # reference: xva_train.py:539-587
    def init_logs(self, dataset_output):
        os.makedirs(dataset_output, exist_ok=True)
        self.training_log = []
        if os.path.exists(f"{dataset_output}/training.log"):
            with open(f"{dataset_output}/training.log") as f:
                self.training_log = [l for l in f.read().split("\n") if l]
        gpath = f"{dataset_output}/graphs.json"
        if os.path.exists(gpath):
            self.graphs_json = json.load(open(gpath))
        else:
            self.graphs_json = {"stages": {str(s): {"loss": [], "loss_delta": [], "target_delta": None} for s in range(1, 6)}}
        self.logs_are_init = True

    def _write_graphs(self):
        if self.rank == 0:
            with open(f"{self.dataset_output}/graphs.json", "w+") as f:
                f.write(json.dumps(self.graphs_json))

    def _scalars(self, step, **named):
        """TensorBoard scalars under the output folder, rank 0 only (SummaryWriter(log_dir=dataset_output, flush_secs=120),
        xva_train.py:297; tags as xva_train.py:841-848,892-899,944-948 and hifigan/xva_train.py:554-557,628)."""
        if self.rank != 0:
            return
        if getattr(self, "_tb", None) is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(log_dir=self.dataset_output, flush_secs=120)
            except Exception:
                self._tb = False
        if self._tb:
            for tag, v in named.items():
                if v is not None:
                    self._tb.add_scalar(tag, float(v), int(step))

    def _dist_init(self):
        """Join the NCCL process group when launched under torchrun (idempotent). -> world size"""
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.device)
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=self.device)
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
        return self.world

    async def _send(self, msg):
        if self.websocket is not None and self.rank == 0:
            r = self.websocket.send(msg)
            if asyncio.iscoroutine(r):
                await r


# ==================================================================================================== FastPitch
class FastPitchTrainer(_TrainerBase):
    """python/fastpitch1_1/xva_train.py:185. All four stages run on the B200 engine. A run starts at stage 1 (the
    aligner) when its batches carry the alignment prior (x[7]), otherwise at stage 2 with the durations they carry;
    when stage 1 ends, the durations of every batch are replaced by the aligner's (the in-memory equivalent of the
    reference's duration extraction to durs_arpabet/*.npy, xva_train.py:1128-1160)."""

    KL_LOSS_START_EPOCH, KL_LOSS_WARMUP_EPOCHS, KL_LOSS_WEIGHT = 0, 100, 1.0    # xva_train.py:706-708

    def __init__(self, logger, PROD, gpus, models_manager, websocket=None):
        super().__init__(logger, PROD, gpus, models_manager, websocket)
        if logger:
            logger.info("New FastPitchTrainer")

    async def start(self, data, gpus=None, resume=False):
        if self.running:
            return
        self.running = True
        if not resume:
            self.force_stage = int(data["force_stage"]) if data.get("force_stage") else None
            self.dataset_input = data["dataset_path"]
            self.dataset_id = (self.dataset_input.split("/")[-1] or "voice").replace(":", "_")
            self.dataset_output = os.path.join(data["output_path"], self.dataset_id)
            self.checkpoint = data.get("checkpoint")
            self.workers = data.get("num_workers", 0)
            self.batch_size = int(data["batch_size"])
            self.epochs_per_checkpoint = int(data.get("epochs_per_checkpoint", 1))
            self.batch_source = data.get("batch_source")
            self.learning_rate, self.warmup_steps = 0.1, 1000
            self.max_epochs = int(os.environ.get("XVA_B200_MAX_EPOCHS", "0"))       # test hook: bound epochs per stage
        if not self.logs_are_init:
            self.init_logs(self.dataset_output)
        while self.running:
            await self.iteration()

    def _batches_key_changed(self, key):
        return getattr(self, "_batches_key", None) != key

    async def init(self):
        world = self._dist_init()
        torch.manual_seed(1234 + self.rank)
        self.model = fp.FastPitch(logger=self.logger, device=self.device, seed=1234)   # same initial weights on every rank
        self.model.seed = 1234 + self.rank                                         # per-rank dropout streams (:294-295)
        self.criterion = fp.FastPitchLoss().set_distributed(world)
        self.attention_kl_loss = fp.AttentionBinarizationLoss().set_distributed(world)   # :342
        self.optimizer = fp.Lamb(self.model, lr=0.1, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        # gradient all-reduce overlapped with backward; losses are normalised by GLOBAL mask sums, so SUM is exact (8e)
        self.sync = parallel.GradSync(self.model, world, mean=False) if world > 1 else None
        self.total_iter, self.epoch, self.avg_loss_per_epoch = 50000, 0, []       # new voices start at 40-50k (:304-335)
        self.start_iterations = self.total_iter
        stage = None
        # xva_train.py:284-288: the newest checkpoint of THIS run wins; the user's base checkpoint is only the starting point
        # of a run whose output folder has none yet (handleTrainer re-enters init() after every stage change and on resume).
        ck = self.last_checkpoint(self.dataset_output)
        if ck is None and self.checkpoint and os.path.isfile(str(self.checkpoint)):
            ck = self.checkpoint
            self.print_and_log(f"Checkpoint: {ck}", save_to_file=self.dataset_output)
        loaded_stage = None
        if ck:
            loaded_stage, self.epoch, self.total_iter, self.avg_loss_per_epoch = self.load_checkpoint(ck)
            self.ckpt_path = ck
            stage = loaded_stage
            # :356-359: a finished run (stage 5 on disk, or stage 5 forced) goes straight on to the vocoder
            if (int(loaded_stage) == 5 and self.force_stage is None) or self.force_stage == 5:
                self.END_OF_TRAINING = True
                self.JUST_FINISHED_STAGE = True
                raise StageFinished("FastPitch already finished: move to HiFi-GAN")
        elif self.force_stage == 5:
            self.END_OF_TRAINING = True
            self.JUST_FINISHED_STAGE = True
            raise StageFinished("stage 5 forced: move to HiFi-GAN")
        if self.force_stage:                                                      # :361-366
            self.print_and_log(f"Forcing stage: {self.force_stage}", save_to_file=self.dataset_output)
            if loaded_stage is not None and int(loaded_stage) < self.force_stage and int(loaded_stage) != 3:
                self.total_iter = self.start_iterations
            self.avg_loss_per_epoch = []
            stage = self.force_stage
        if ck and self.dataset_id not in str(ck):                                 # IS_NEW, :380-385: a base checkpoint of another voice
            self.print_and_log("New voice", save_to_file=self.dataset_output)
            stage = None                   # decided below: 1 when the batches carry the alignment prior (the reference's 1)
            self.total_iter = self.start_iterations
            self.avg_loss_per_epoch = []
            self.epoch = 0
        source_key = (str(self.dataset_input), id(self.batch_source))
        if getattr(self, "_batches_key", None) == source_key:
            pass          # same trainer re-initialised for the next stage: keep the batches (and the extracted durations)
        elif self.batch_source is not None:
            self.batches = parallel.shard_batches(self.batch_source, self.rank, world)    # rank r: items r::W, drop_last
        elif str(self.dataset_input).startswith("synthetic:"):
            self.batches = _synthetic_fastpitch_batches(self.dataset_input, self.device, 1234 + self.rank)
        else:
            raise NotImplementedError("wav/text dataset loading is outside this build (SURVEY.md section 2 row 6): pass "
                                      "data['batch_source'] or dataset_path='synthetic:BxTtxTmxitems'")
        if self._batches_key_changed(source_key):
            self.batches = parallel.pad_batches_to_global(self.batches, world)     # DataParallel pads the global batch (8e)
        self._batches_key = source_key
        has_prior = all(b[0][7] is not None for b in self.batches)
        if stage is None:
            stage = 1 if has_prior else 2
        stage = int(stage)
        if stage not in (1, 2, 3, 4):
            raise ValueError(f"training stage {stage} is not a FastPitch stage (1-4)")
        if stage == 1 and not has_prior:
            raise ValueError("training stage 1 needs batches that carry the alignment prior (inputs_x[7])")
        self.model.training_stage = self.criterion.training_stage = stage
        num_data_lines = world * sum(int(b[0][0].shape[0]) for b in self.batches)   # utterances in the dataset (:380)
        self.target_delta = get_target_delta(stage, num_data_lines)
        self.graphs_json["stages"][str(stage)]["target_delta"] = self.target_delta
        await self._send(f"Set stage to: {stage} ")
        mult = {1: 1.5, 2: 12, 3: 3.5, 4: 4}.get(stage, 1)                         # stage batch multipliers (:387-404)
        self.stage_batch = max(1, int(self.batch_size * mult))
        self.gam = max(1, round(256 / (self.stage_batch * world)))                 # :403-407: global batch = W x per-GPU batch
        if stage == 1:
            self.print_and_log("Stage 1: Pre-training only the alignment.", save_to_file=self.dataset_output)   # :411-412
        self.model.train()
        self.EPOCH_AVG_SPAN, self.target_patience, self.target_patience_count = 20, 3, 0
        self.last_loss, self.iter_losses, self.avg_frames_s = None, [], []
        self.epoch_iter, self.micro, self.frames_acc = 0, 0, 0
        self._window_loss = None
        self.use_graph = os.environ.get("XVA_TRAINER_GRAPH", "1") != "0"
        self._graphs, self._lens_cache, self._graph_pool = {}, {}, None    # per stage: init() runs again after a stage change
        self.optimizer.lr_on_device = False
        self.batch_pos = 0
        self.avg_loss_per_epoch.append(0.0)
        self.step_t0 = time.perf_counter()
        self.model.zero_grad()
        self.is_init = True

    # ---- CUDA-graph replay of the micro-step (stages 2-4). One graph per (stage, batch shape, last-of-window): the step is
    # ~330 launches with static shapes, and issued one by one from Python it is host-bound (bench: 14.9 ms eager vs 13.7 ms
    # replayed at 32 x 880 x 160; far worse at small batches). Everything that changes between replays lives on the device:
    # the batch (copied into the graph's static input tensors), the learning rate (Lamb.lr_dev), the dropout counter, and the
    # skip-on-non-finite decision (xva_lamb_step is a no-op when the gradient norm is not finite). Stage 1 stays eager: its
    # binarization-loss weight changes every epoch and is a kernel argument.
    GRAPH_CACHE = 8

    def _host_lens(self, x):
        """(mel_max_len, max decoder length) of a batch as Python ints, read once per batch and cached: FastPitch.forward
        then needs no device->host read (model.py:330 and :75 make two per step)."""
        key = (id(x), id(x[8]))
        hit = self._lens_cache.get(key)
        if hit is None:
            mel_max = int(x[10][0].item())
            _, dec = ops.duration_scan(x[8].to(torch.float32), 1.0, mel_max)
            both = torch.stack([torch.tensor(mel_max, device=dec.device, dtype=torch.int64), dec.max().to(torch.int64)])
            if self.world > 1:
                # every rank runs its decoder at the GLOBAL maximum length, as the replicas of nn.DataParallel do: an
                # utterance's last frames depend on the padded length of its batch (parallel.pad_batches_to_global)
                import torch.distributed as dist
                dist.all_reduce(both, op=dist.ReduceOp.MAX)
            hit = self._lens_cache[key] = tuple(int(v) for v in both.tolist())
        return hit

    def _micro_step(self, x, y, last, host_lens):
        """forward + loss + backward of one micro-batch; with ``last`` also gradient exchange, LAMB, dropout counter and
        zero_grad. Returns device scalars only (no host sync): tracked loss, gradient norm, the logged loss terms."""
        stage = self.model.training_stage
        y_pred = self.model(x, host_lens=host_lens)                                                          # :788
        loss, meta = self.criterion(y_pred, y)                                                               # :790
        sync = self.sync if last else None
        self.model.backward(self.criterion, 1.0 / self.gam, grad_sync=sync)                                  # :806-813
        if sync is not None:
            sync.finish()
        key = {3: "pitch_loss", 4: "mel_loss"}.get(stage, "loss")                                            # :815-820
        tracked = (meta[key] * (0.1 if stage == 3 else 1.0)).double().reshape(())
        names = ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss")
        zero = torch.zeros((), device=tracked.device, dtype=torch.float64)
        terms = [meta[k].double().reshape(()) if torch.is_tensor(meta.get(k)) else zero for k in names]
        gsq = zero
        if last:
            gsq = torch.linalg.vector_norm(self.model.arena.g).double().reshape(())     # NaN / Inf propagate
            self.optimizer.step()          # no-op on the device when the gradient norm is not finite (xva_lamb_step)      :855-862
            self.model.step_dropout()
            self.model.zero_grad()
        return torch.stack([tracked, gsq] + terms)

    def _graphed_micro_step(self, x, y, last):
        stage = self.model.training_stage
        hl = self._host_lens(x)
        tensors = [i for i, t in enumerate(x) if torch.is_tensor(t)]
        sig = (stage, bool(last), hl, tuple((i, tuple(x[i].shape), x[i].dtype) for i in tensors))
        entry = self._graphs.get(sig)
        if entry is None:
            if len(self._graphs) >= self.GRAPH_CACHE:
                return None                                   # shape not worth another graph: run it eagerly
            static_x = [t.clone() if torch.is_tensor(t) else t for t in x]
            static_y = [static_x[2], static_x[1], static_x[3], static_x[9]]
            self.optimizer.lr_on_device = True

            def run(*ts):
                return self._micro_step(static_x, static_y, last, hl)

            # warm-up executions inside GraphedStep would apply real optimizer steps: capture without warm-up, after one
            # eager execution of this very micro-step (the caller's), so every lazily created buffer exists
            entry = self._graphs[sig] = (None, static_x, tensors)
            return None
        gstep, static_x, tensors = entry
        if gstep is None:
            static_y = [static_x[2], static_x[1], static_x[3], static_x[9]]
            self.optimizer.lr_on_device = True               # the captured LAMB reads lr_dev, filled before every replay
            gstep = graph.GraphedStep(lambda: self._micro_step(static_x, static_y, last, hl), [], warmup=0, pool=self._graph_pool)
            self._graph_pool = gstep.pool
            self._graphs[sig] = (gstep, static_x, tensors)
            # the capture itself does not execute: fall through to the replay below
        for i in tensors:
            if static_x[i] is not x[i]:
                static_x[i].copy_(x[i], non_blocking=True)
        self.optimizer.lr_dev.fill_(float(self.optimizer.param_groups[0]["lr"]))
        return gstep()

    async def iteration(self):
        if not self.is_init:
            await self.init()
        if self.batch_pos >= len(self.batches):
            self.batch_pos = 0
            self.finish_epoch()
            self.avg_loss_per_epoch.append(0.0)
            self.epoch_iter = 0
            self.iter_losses = []
        if self.use_graph and self.model.training_stage != 1:
            return await self._iteration_graphed()
        x, y, num_frames = self.batches[self.batch_pos]
        self.batch_pos += 1
        self.total_iter += 1
        self.epoch_iter += 1
        fp.adjust_learning_rate(self.total_iter, self.optimizer, self.learning_rate, self.warmup_steps)     # :780
        stage = self.model.training_stage
        y_pred = self.model(x) if stage == 1 else self.model(x, host_lens=self._host_lens(x))                # :788
        loss, meta = self.criterion(y_pred, y)                                                               # :790
        kl = None
        if stage == 1 and self.KL_LOSS_START_EPOCH is not None and self.epoch >= self.KL_LOSS_START_EPOCH:   # :792-798
            binarization_loss = self.attention_kl_loss(y_pred[9], y_pred[8])
            kl_weight = min((self.epoch - self.KL_LOSS_START_EPOCH) / self.KL_LOSS_WARMUP_EPOCHS, 1.0) * self.KL_LOSS_WEIGHT
            meta["kl_loss"] = binarization_loss * kl_weight
            meta["loss"] = meta["loss"] + kl_weight * binarization_loss
            kl = (self.attention_kl_loss, kl_weight)
        self.micro += 1
        # the gradient exchange rides on the backward of the window's LAST micro-batch (the arena then holds the sum)
        sync = self.sync if (self.sync is not None and self.micro % self.gam == 0) else None
        self.model.backward(self.criterion, 1.0 / self.gam, grad_sync=sync, kl=kl)                           # :806-813
        if sync is not None:
            sync.finish()
        self.frames_acc += num_frames
        key = {3: "pitch_loss", 4: "mel_loss"}.get(stage, "loss")                                            # :815-820
        tracked = meta[key] * (0.1 if stage == 3 else 1.0)
        if self.micro % self.gam != 0:
            self._window_loss = tracked if self._window_loss is None else self._window_loss + tracked
        if self.micro % self.gam == 0:
            # NaN / Inf guard BEFORE the update (xva_train.py:825-832 checks every micro-batch and skips): one host read per
            # optimizer step of the tracked loss summed over the window and of the squared gradient norm of the arena, so
            # a non-finite micro-batch anywhere in the accumulation window is seen. On a hit nothing is applied: gradients
            # are cleared, moments, weights and their tf32 copy stay as they were.
            self._window_loss = tracked if self._window_loss is None else self._window_loss + tracked
            gsq = torch.linalg.vector_norm(self.model.arena.g).double()     # NaN / Inf propagate
            names = ("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss", "kl_loss")
            zero = torch.zeros((), device=gsq.device, dtype=torch.float64)
            terms = [meta[k].double().reshape(()) if k in meta else zero for k in names]
            host = torch.stack([self._window_loss.double().reshape(()), gsq.reshape(())] + terms).tolist()
            val, gval, logged = float(host[0]) / self.gam, float(host[1]), dict(zip(names, host[2:]))
            self._window_loss = None
            if not (math.isfinite(val) and math.isfinite(gval)):
                self.model.zero_grad()
                self.frames_acc = 0
                self.step_t0 = time.perf_counter()
                self.print_and_log("loss is NaN", save_to_file=self.dataset_output)
                return
            self.optimizer.step()                                                                            # :855-862
            self.model.step_dropout()
            self.model.zero_grad()
            dt = time.perf_counter() - self.step_t0
            self.step_t0 = time.perf_counter()
            frames_s = self.world * self.frames_acc / max(dt, 1e-9)        # whole job (every rank holds the same frame count)
            self.frames_acc = 0
            self.avg_frames_s.append(frames_s)
            self.iter_losses.append(val)
            self.avg_loss_per_epoch[-1] += val
            nz = lambda v: v if v else None             # the reference only writes the terms the stage trains (:839-848)
            self._scalars(self.total_iter, **{"loss/loss": logged["loss"], "loss/mel": nz(logged["mel_loss"]),
                                              "loss/dur": nz(logged["duration_predictor_loss"]), "loss/pitch": nz(logged["pitch_loss"]),
                                              "loss/energy": nz(logged["energy_loss"]), "loss/kl": nz(logged["kl_loss"]),
                                              "meta/frames/s": frames_s, "meta/lrate": self.optimizer.param_groups[0]["lr"]})
            self.training_log_live_line = (f"| Stage {stage} | Epoch {self.epoch} | iter {self.total_iter} | loss "
                                           f"{val:.5f} | frames/s {int(frames_s)} | lr {self.optimizer.param_groups[0]['lr']:.2e}")
            self.print_and_log(save_to_file=self.dataset_output)

    async def _iteration_graphed(self):
        """iteration() for stages 2-4 with the micro-step replayed from a CUDA graph (eager for the first two sightings of
        a shape: one to create every lazily allocated buffer, one inside the capture's own bookkeeping)."""
        x, y, num_frames = self.batches[self.batch_pos]
        self.batch_pos += 1
        self.total_iter += 1
        self.epoch_iter += 1
        fp.adjust_learning_rate(self.total_iter, self.optimizer, self.learning_rate, self.warmup_steps)     # :780
        stage = self.model.training_stage
        self.micro += 1
        last = self.micro % self.gam == 0
        steps_before = self.optimizer.steps      # host-side count of APPLIED updates (a capture or a replay does not run
        out = self._graphed_micro_step(x, y, last)   # Lamb.step's Python; a skipped update must not count)
        if out is None:
            self.optimizer.lr_on_device = False
            out = self._micro_step(x, y, last, self._host_lens(x))
        self.optimizer.steps = steps_before
        self.frames_acc += num_frames
        # (out is the graph's static output tensor: the next replay overwrites it)
        self._window_loss = out[0].clone() if self._window_loss is None else self._window_loss + out[0]
        if not last:
            return
        host = torch.cat([self._window_loss.reshape(1), out[1:]]).tolist()          # the one host read of the optimizer step
        self._window_loss = None
        val, gval = float(host[0]) / self.gam, float(host[1])
        logged = dict(zip(("loss", "mel_loss", "duration_predictor_loss", "pitch_loss", "energy_loss"), host[2:]))
        if not (math.isfinite(val) and math.isfinite(gval)):
            # the device already skipped the update (non-finite gradient norm); a NaN loss with finite gradients cannot
            # happen (the loss is a function of the same activations), so there is nothing to undo
            self.frames_acc = 0
            self.step_t0 = time.perf_counter()
            self.print_and_log("loss is NaN", save_to_file=self.dataset_output)
            return
        self.optimizer.steps = steps_before + 1
        dt = time.perf_counter() - self.step_t0
        self.step_t0 = time.perf_counter()
        frames_s = self.world * self.frames_acc / max(dt, 1e-9)
        self.frames_acc = 0
        self.avg_frames_s.append(frames_s)
        self.iter_losses.append(val)
        self.avg_loss_per_epoch[-1] += val
        nz = lambda v: v if v else None
        self._scalars(self.total_iter, **{"loss/loss": logged["loss"], "loss/mel": nz(logged["mel_loss"]),
                                          "loss/dur": nz(logged["duration_predictor_loss"]), "loss/pitch": nz(logged["pitch_loss"]),
                                          "loss/energy": nz(logged["energy_loss"]),
                                          "meta/frames/s": frames_s, "meta/lrate": self.optimizer.param_groups[0]["lr"]})
        self.training_log_live_line = (f"| Stage {stage} | Epoch {self.epoch} | iter {self.total_iter} | loss "
                                       f"{val:.5f} | frames/s {int(frames_s)} | lr {self.optimizer.param_groups[0]['lr']:.2e}")
        self.print_and_log(save_to_file=self.dataset_output)

    # reference: xva_train.py:915-976
    def finish_epoch(self):
        self.epoch += 1
        n_steps = max(1, len(self.iter_losses))
        self.avg_loss_per_epoch[-1] /= n_steps
        deltas = [(a - b) / a for a, b in zip(self.avg_loss_per_epoch[:-1], self.avg_loss_per_epoch[1:]) if a]
        avg_loss = sum(self.iter_losses) / n_steps if self.iter_losses else 0.0
        delta_avg = None
        if len(deltas) >= 2:
            span = deltas[-self.EPOCH_AVG_SPAN:]
            delta_avg = sum(span) / len(span)
        stage = self.model.training_stage
        fpath = os.path.join(self.dataset_output, f"FastPitch_checkpoint_{self.epoch}_{self.total_iter}.pt")
        frames_s = sum(self.avg_frames_s) / max(1, len(self.avg_frames_s))
        self.save_checkpoint(False, frames_s, self.total_iter, avg_loss, delta_avg, self.avg_loss_per_epoch, fpath)
        self.graphs_json["stages"][str(stage)]["loss"].append([self.total_iter, self.avg_loss_per_epoch[-1]])
        if deltas:
            self._scalars(self.total_iter, **{"meta/acc_epoch_delta": deltas[-1]})
        if delta_avg is not None:
            self._scalars(self.total_iter, **{f"meta/stage_{stage}_acc_epoch_deltas_avg20": delta_avg})
            self.graphs_json["stages"][str(stage)]["loss_delta"].append([self.total_iter, delta_avg])
        self._write_graphs()
        done = False
        if delta_avg is not None and len(deltas) >= (20 if stage == 2 else 1) and delta_avg <= self.target_delta:
            self.target_patience_count += 1
            done = self.target_patience_count >= self.target_patience
        else:
            self.target_patience_count = 0
        if self.max_epochs and self.epoch_in_stage() >= self.max_epochs:
            done = True
        self.avg_frames_s = []
        if done:
            stage_path = os.path.join(self.dataset_output, f"Stage_{stage}_DONE_FastPitch_checkpoint_{self.epoch}_{self.total_iter}.pt")
            if stage == 4:
                self.END_OF_TRAINING = True
            self.JUST_FINISHED_STAGE = True
            if stage == 1:
                self.extract_durations()
            self.model.training_stage += 1
            self.avg_loss_per_epoch = []
            it = self.total_iter if self.model.training_stage == 4 else self.start_iterations
            self.save_checkpoint(True, frames_s, it, avg_loss, delta_avg, self.avg_loss_per_epoch, fpath)
            self.save_checkpoint(True, frames_s, it, avg_loss, delta_avg, self.avg_loss_per_epoch, stage_path, doPrintLog=False)
            raise StageFinished(f"stage {stage} finished")

    def epoch_in_stage(self):
        return len(self.avg_loss_per_epoch)

    def extract_durations(self):
        """End of stage 1 (xva_train.py:1128-1160): run the aligner over every batch without gradients and keep its hard
        durations as the duration targets of stages 2-4. The reference writes them to durs_arpabet / durs_text .npy
        files that its dataset reads back; batches here are in memory, so x[8] is replaced in place."""
        self.print_and_log("Extracting durations from alignments...", save_to_file=self.dataset_output)
        was_training = self.model.training
        self.model.eval()
        with torch.no_grad():
            for x, _, _ in self.batches:
                x[8] = self.model(x)[10]
        self.model.train(was_training)

    # reference: xva_train.py:979-1052
    def save_checkpoint(self, force_save=False, frames_s=0, total_iter=0, avg_loss=None, loss_delta=None,
                        avg_loss_per_epoch=(), fpath="out.pt", doPrintLog=True):
        if self.rank != 0:
            return
        intermediate = self.epochs_per_checkpoint > 0 and self.epoch % self.epochs_per_checkpoint == 0
        if not intermediate and not force_save:
            return
        old = sorted([f for f in os.listdir(self.dataset_output) if f.startswith("FastPitch_checkpoint_")], key=_sort_fp)
        for f in old[:-2] if len(old) > 2 else []:
            os.remove(f"{self.dataset_output}/{f}")
        line = (f"Stage: {self.model.training_stage} | Epoch: {self.epoch} | {os.path.basename(self.dataset_output)}~"
                f"{self.epoch}_{self.total_iter}.pt | frames/s: {int(frames_s)}")
        if avg_loss is not None:
            line += f" | Loss: {avg_loss:.5f}"
        if loss_delta is not None:
            line += f" | Delta: {loss_delta:.5f}"
        line += f" | Target: {self.target_delta:.5f}"
        sd = self.model.state_dict()
        torch.save({"epoch": self.epoch, "iteration": total_iter, "avg_loss_per_epoch": list(avg_loss_per_epoch),
                    "training_stage": self.model.training_stage, "state_dict": sd,
                    "optimizer": self.optimizer.state_dict()}, fpath)
        # the fp16 export xVASynth loads (:1013-1016) is made from a copy: the live fp32 model is never touched
        torch.save({k: (v.half() if v.is_floating_point() else v) for k, v in sd.items()},
                   f"{self.dataset_output}/{self.dataset_id}.pt")
        with open(f"{self.dataset_output}/{self.dataset_id}.json", "w+") as f:
            json.dump({"version": "2.0", "modelVersion": "2.0", "modelType": "FastPitch1.1", "author": "", "lang": "en",
                       "games": [{"gameId": "other", "voiceId": self.dataset_id,
                                  "voiceName": os.path.basename(self.dataset_output), "resemblyzer": [], "gender": "male"}]},
                      f, indent=4)
        self.training_log_live_line = ""
        if doPrintLog:
            self.print_and_log(line, save_to_file=self.dataset_output)

    # reference: xva_train.py:1054-1081
    def load_checkpoint(self, filepath):
        self.print_and_log(f"Loading model and optimizer state from {filepath}", save_to_file=self.dataset_output)
        try:
            ck = torch.load(filepath, map_location="cpu")
        except Exception:
            self.print_and_log("Failed to load the checkpoint! Maybe try the second-last checkpoint (delete the last one). "
                               f"Full error message: {traceback.format_exc()}", save_to_file=self.dataset_output)
            raise
        sd = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
        self.model.load_state_dict(sd)
        try:
            self.optimizer.load_state_dict(ck["optimizer"])
        except Exception:
            self.print_and_log("========== OPTIM NOT LOADED ==========", save_to_file=self.dataset_output)
        return ck.get("training_stage", 1), ck.get("epoch", 0) + 1, ck.get("iteration", 0), ck.get("avg_loss_per_epoch", [])

    @staticmethod
    def last_checkpoint(output):      # :1239-1250
        if not output or not os.path.isdir(output):
            return None
        c = sorted([f for f in os.listdir(output) if f.startswith("FastPitch_checkpoint_")], key=_sort_fp)
        return os.path.join(output, c[-1]) if c else None


def _sort_fp(name):
    return int(name.split("FastPitch_checkpoint_")[-1].split(".")[0].split("_")[0])


async def handleTrainer(models_manager, data, websocket, gpus, resume=False):
    """python/fastpitch1_1/xva_train.py:57-176: drives FastPitch stages 1 -> 4, returns "move to hifi" when stage 4 ends."""
    gpus = gpus or [0]
    trainer = models_manager.sync_init_model("fastpitch1_1", websocket=websocket, gpus=gpus)
    try:
        await trainer.start(data, gpus=gpus, resume=resume)
        return None
    except RuntimeError as e:
        if "out of memory" in str(e):                                   # :131-145: retry three items smaller
            trainer.print_and_log(f"Out of VRAM; batch size {data['batch_size']} -> {int(data['batch_size']) - 3}")
            data["batch_size"] = max(1, int(data["batch_size"]) - 3)
            trainer.running, trainer.is_init = False, False
            torch.cuda.empty_cache()
            return await handleTrainer(models_manager, data, websocket, gpus)
        if trainer.JUST_FINISHED_STAGE:                                 # :147-168
            trainer.JUST_FINISHED_STAGE = False
            trainer.running, trainer.is_init = False, False
            if trainer.END_OF_TRAINING:
                trainer.END_OF_TRAINING = False
                del models_manager.models_bank["fastpitch1_1"]
                return "move to hifi"
            data["force_stage"] = None
            return await handleTrainer(models_manager, data, websocket, gpus)
        raise


# ==================================================================================================== HiFi-GAN
class _Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


HIFI_CONFIG_V1 = dict(resblock="1", num_gpus=0, batch_size=16, learning_rate=0.0002, adam_b1=0.8, adam_b2=0.99,
                      lr_decay=0.999, seed=1234, upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                      upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                      resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], segment_size=8192, num_mels=80, num_freq=1025,
                      n_fft=1024, hop_size=256, win_size=1024, sampling_rate=22050, fmin=0, fmax=8000, fmax_for_loss=None)


class HiFiTrainer(_TrainerBase):
    """python/hifigan/xva_train.py:132 ("stage 5")."""

    def __init__(self, logger, PROD, gpus, models_manager, websocket=None):
        super().__init__(logger, PROD, gpus, models_manager, websocket)
        if logger:
            logger.info("New HiFiTrainer")

    async def start(self, data, gpus=None, resume=False):
        if self.running:
            return
        self.running = True
        if not resume:
            self.dataset_input = data["dataset_path"]
            self.dataset_id = (self.dataset_input.split("/")[-1] or "voice").replace(":", "_")
            self.dataset_output = os.path.join(data["output_path"], self.dataset_id)
            self.hifi_dir = os.path.join(self.dataset_output, "hifi")
            self.checkpoint = data.get("hifigan_checkpoint")
            self.batch_size = int(data["batch_size"])
            self.epochs_per_checkpoint = int(data.get("epochs_per_checkpoint", 1))
            self.batch_source = data.get("batch_source")
            self.max_epochs = int(os.environ.get("XVA_B200_MAX_EPOCHS", "0"))
        if not self.logs_are_init:
            self.init_logs(self.dataset_output)
        while self.running:
            await self.iteration()

    async def init(self):
        os.makedirs(self.hifi_dir, exist_ok=True)
        h = _Cfg(HIFI_CONFIG_V1)
        h.batch_size = int(self.batch_size * 1.4)                                     # :228
        self.h = h
        world = self._dist_init()
        torch.manual_seed(h.seed + self.rank)
        self.generator = hg.Generator(h, device=self.device)
        self.mpd = hg.MultiPeriodDiscriminator(device=self.device)
        self.msd = hg.MultiScaleDiscriminator(device=self.device)
        self.model = self.generator
        self.steps, self.epoch, self.avg_loss_per_epoch = 0, 0, []
        cp_g, cp_do = self.scan_checkpoint("g_"), self.scan_checkpoint("do_")
        if self.checkpoint and os.path.isfile(str(self.checkpoint)):
            cp_g = self.checkpoint
        synthetic = self.batch_source is not None or str(self.dataset_input).startswith("synthetic:")
        if cp_g is None and not synthetic:
            raise RuntimeError("HiFi-GAN fine-tuning needs a generator checkpoint (hifigan/xva_train.py:276-277)")
        if cp_g:
            self.generator.load_state_dict(torch.load(cp_g, map_location="cpu")["generator"])
            self.ckpt_path = cp_g
        state_do = torch.load(cp_do, map_location="cpu") if cp_do else None
        if state_do:
            self.mpd.load_state_dict(state_do["mpd"])
            self.msd.load_state_dict(state_do["msd"])
            self.steps, self.epoch = state_do["steps"] + 1, state_do["epoch"]
            self.avg_loss_per_epoch = list(state_do.get("avg_loss_per_epoch", []))
        self.generator.train(); self.mpd.train(); self.msd.train()
        self.stepper = hg.HiFiGANStep(self.generator, self.mpd, self.msd, h, world=world)
        if state_do and "optim_g" in state_do:
            for opt, key in ((self.stepper.optim_g, "optim_g"), (self.stepper.optim_d, "optim_d")):
                try:
                    opt.load_state_dict(state_do[key])
                except Exception:
                    self.print_and_log(f"========== OPTIM NOT LOADED ({key}) ==========", save_to_file=self.dataset_output)
        self.lr = h.learning_rate * (h.lr_decay ** self.epoch)                        # ExponentialLR per epoch (:306-307)
        self.wav_segments = None
        if self.batch_source is not None:
            self.batches = parallel.shard_batches(self.batch_source, self.rank, world)
        elif str(self.dataset_input).startswith("synthetic:"):
            B, frames, items = (int(v) for v in self.dataset_input.split(":", 1)[1].split("x"))
            from . import hifigan as _hg
            g = torch.Generator().manual_seed(1234 + self.rank)
            mel_in = _hg.MelSpectrogram(fmax=h.fmax, device=self.device)
            mel_loss = _hg.MelSpectrogram(fmax=h.fmax_for_loss, device=self.device)
            self.batches = []
            for _ in range(max(1, items // B)):
                y = (0.95 * torch.tanh(torch.randn(B, frames * h.hop_size, generator=g) * 0.3)).to(self.device)
                self.batches.append((mel_in(y).transpose(1, 2).contiguous(), y, mel_loss(y).transpose(1, 2).contiguous()))
        elif os.path.isfile(os.path.join(str(self.dataset_input), "metadata.csv")):
            # a voice folder (metadata.csv + wavs/): hifigan/xva_train.py:309-325 with the crops assembled on the host and
            # both mel spectrograms computed per batch on the device (wavdata.py)
            from . import hifigan as _hg, wavdata
            files, not_found, dm = wavdata.get_dataset_filelist(f"{self.dataset_input}/metadata.csv", f"{self.dataset_input}/wavs")
            self.print_and_log(f"Training items: {int(len(files) / dm)} | Data multiplier: {dm} | Not found: {not_found} | "
                               f"Total: {len(files)}", save_to_file=self.dataset_output)                      # :313
            self.wav_segments = wavdata.WavSegments(files, h.segment_size, h.sampling_rate)
            self.mel_extractors = (_hg.MelSpectrogram(fmax=h.fmax, device=self.device),
                                   _hg.MelSpectrogram(fmax=h.fmax_for_loss, device=self.device))
            self.batches = self._wav_epoch()
        else:
            raise NotImplementedError("dataset_path must be a voice folder (metadata.csv + wavs/), 'synthetic:BxFRAMESxITEMS', "
                                      "or pass data['batch_source']")
        self.target_delta, self.target_patience, self.target_patience_count = 0.0001, 3, 0     # :268-271
        self.EPOCH_AVG_SPAN = 20
        self.graphs_json["stages"]["5"]["target_delta"] = self.target_delta
        await self._send("Set stage to: 5 ")
        self.batch_pos, self.iter_losses = 0, []
        self.avg_loss_per_epoch.append(0.0)
        self.use_graph = os.environ.get("XVA_TRAINER_GRAPH", "1") != "0"
        self._gstep, self._gstep_sig = None, None
        self.is_init = True

    def _wav_epoch(self):
        """One epoch of batches from the voice folder: fresh permutation, fresh random crops (meldataset.py:358-362)."""
        bs = int(self.h.batch_size)                                                   # DataLoader(batch_size=h.batch_size), :321
        batches = list(self.wav_segments.batches(bs, self.device, *self.mel_extractors))
        if not batches:
            raise ValueError(f"batch size {bs} exceeds the {len(self.wav_segments)} training items")
        return batches

    async def iteration(self):
        if not self.is_init:
            await self.init()
        if self.batch_pos >= len(self.batches):
            self.batch_pos = 0
            self.finish_epoch()
            self.avg_loss_per_epoch.append(0.0)
            self.iter_losses = []
            if getattr(self, "wav_segments", None) is not None:
                self.batches = self._wav_epoch()
        x, y, y_mel = self.batches[self.batch_pos]
        self.batch_pos += 1
        t0 = time.perf_counter()
        opts = (self.stepper.optim_g, self.stepper.optim_d)
        for opt in opts:
            opt.param_groups[0]["lr"] = self.lr
        sig = (tuple(x.shape), tuple(y.shape), tuple(y_mel.shape))
        if self.use_graph and self._gstep is not None and self._gstep_sig == sig:
            # one CUDA-graph replay per step (~1 000 launches; eager the step is host-bound: 41.7 vs 32.1 ms at 16 x 8192);
            # learning rate and AdamW step count live on the device
            for opt in opts:
                opt.lr_dev.fill_(float(self.lr))
            out = self._gstep(x, y, y_mel)
            for opt in opts:
                opt.steps += 1
            self.stepper.steps += 1
        else:
            for opt in opts:
                opt.lr_on_device = False
            out = self.stepper.step(x, y, y_mel)                                      # :467-515
            if self.use_graph and self._gstep is None:
                # captured after this first eager step (every lazily created buffer now exists); capture does not execute
                for opt in opts:
                    opt.lr_on_device = True
                statics = [x.clone(), y.clone(), y_mel.clone()]
                self._gstep = graph.GraphedStep(lambda a, b, c: self.stepper.step(a, b, c), statics, warmup=0)
                self._gstep_sig = sig
                for opt in opts:          # the capture ran AdamW.step()'s host-side bookkeeping once without a launch
                    opt.steps -= 1
                self.stepper.steps -= 1
        gen_loss, mel_error = (float(v) for v in torch.stack([out["loss_gen_all"].double(), out["mel_error"].double()]).tolist())
        its = 1.0 / max(time.perf_counter() - t0, 1e-9)
        self._scalars(self.steps, **{"training/gen_loss_total": gen_loss, "training/mel_spec_error": mel_error,
                                     "training/d_lr": self.lr, "training/g_lr": self.lr})          # :554-557
        mel_loss = int(mel_error * 1000) / 1000                                       # :518-523: the tracked quantity
        self.iter_losses.append(mel_loss)
        self.avg_loss_per_epoch[-1] += mel_loss
        self.training_log_live_line = (f"| Stage 5 | Epoch {self.epoch} | Steps {self.steps} | Gen loss {gen_loss:.3f} | "
                                       f"Mel err {mel_error:.4f} | {its * y.shape[0]:.1f} its/s")
        self.print_and_log(save_to_file=self.dataset_output)                          # :517-545
        self.steps += 1

    def finish_epoch(self):                                                           # :607-649
        self.epoch += 1
        self.lr *= self.h.lr_decay
        if self.epochs_per_checkpoint > 0 and self.epoch % self.epochs_per_checkpoint == 0:
            self.output_checkpoint()
        self.avg_loss_per_epoch[-1] /= max(1, len(self.iter_losses))
        self.graphs_json["stages"]["5"]["loss"].append([self.steps, self.avg_loss_per_epoch[-1]])
        deltas = [(a - b) / a for a, b in zip(self.avg_loss_per_epoch[:-1], self.avg_loss_per_epoch[1:]) if a]
        done = False
        if len(deltas) >= 2:
            d = sum(deltas[-self.EPOCH_AVG_SPAN:]) / len(deltas[-self.EPOCH_AVG_SPAN:])
            self._scalars(self.steps, **{"meta/stage_5_acc_epoch_deltas_avg20": d})              # :628
            self.graphs_json["stages"]["5"]["loss_delta"].append([self.steps, d])
            # :633-647: converged when the 20-epoch average relative improvement of the mel loss stays at or below
            # 0.0001 for 3 consecutive epochs, with at least 25 deltas on record
            if d <= self.target_delta and len(deltas) >= 25:
                self.target_patience_count += 1
                done = self.target_patience_count >= self.target_patience
            else:
                self.target_patience_count = 0
        self._write_graphs()
        if self.max_epochs and self.epoch >= self.max_epochs:                         # test hook
            done = True
        if done:
            self.training_log_live_line = ""
            self.print_and_log("HiFi-GAN training finished", save_to_file=self.dataset_output)
            self.output_checkpoint()
            self.END_OF_TRAINING = True
            raise StageFinished("HiFi-GAN finished")

    def output_checkpoint(self):                                                      # :570-604
        if self.rank != 0:
            return
        gpath = f"{self.hifi_dir}/g_{self.steps:08d}"
        torch.save({"generator": self.generator.state_dict()}, gpath)
        # optim_g / optim_d in torch.optim.AdamW's own state_dict layout: the reference's do_ files load here and these
        # load there (hifigan/xva_train.py:583-584)
        torch.save({"mpd": self.mpd.state_dict(), "msd": self.msd.state_dict(), "optim_g": self.stepper.optim_g.state_dict(),
                    "optim_d": self.stepper.optim_d.state_dict(), "steps": self.steps, "epoch": self.epoch,
                    "avg_loss_per_epoch": self.avg_loss_per_epoch, "ckpts_finetuned": True}, f"{self.hifi_dir}/do_{self.steps:08d}")
        torch.save({"generator": self.generator.state_dict()}, f"{self.dataset_output}/{self.dataset_id}.hg.pt")
        for prefix in ("g_", "do_"):                                                  # keep the last two (:592-597)
            old = sorted(f for f in os.listdir(self.hifi_dir) if f.startswith(prefix) and len(f) == len(prefix) + 8)
            for f in old[:-2]:
                os.remove(f"{self.hifi_dir}/{f}")
        self.print_and_log(f"Stage: 5 | Epoch: {self.epoch} | Saved {os.path.basename(gpath)}", save_to_file=self.dataset_output)

    def scan_checkpoint(self, prefix):                                                # hifigan/utils.py:57-62
        d = getattr(self, "hifi_dir", None)
        if not d or not os.path.isdir(d):
            return None
        c = sorted(f for f in os.listdir(d) if f.startswith(prefix) and len(f) == len(prefix) + 8)
        return os.path.join(d, c[-1]) if c else None


async def handleTrainerHiFi(models_manager, data, websocket, gpus, resume=False):
    """python/hifigan/xva_train.py:50-125: returns "done" after sending "Finished training HiFi-GAN\\n"."""
    gpus = gpus or [0]
    trainer = models_manager.sync_init_model("hifigan", websocket=websocket, gpus=gpus)
    try:
        await trainer.start(data, gpus=gpus, resume=resume)
        return None
    except RuntimeError as e:
        if "out of memory" in str(e):                                                 # :95-109
            data["batch_size"] = max(1, int(data["batch_size"]) - 3)
            trainer.running, trainer.is_init = False, False
            torch.cuda.empty_cache()
            return await handleTrainerHiFi(models_manager, data, websocket, gpus)
        if trainer.END_OF_TRAINING:
            trainer.running = False
            await trainer._send("Finished training HiFi-GAN\n")                       # :113
            del models_manager.models_bank["hifigan"]
            return "done"
        raise


# ==================================================================================================== registry
class ModelsManager:
    """python/models_manager.py:8-164, trainer part: key -> lazily constructed trainer, owned by ``models_bank``."""

    def __init__(self, logger=None, PROD=False, device="cuda"):
        self.logger, self.PROD, self.device_label, self.models_bank = logger, PROD, device, {}

    def sync_init_model(self, model_key, websocket=None, gpus=(0,)):
        if model_key in self.models_bank and self.models_bank[model_key] != "move to hifi":
            self.models_bank[model_key].websocket = websocket
            return self.models_bank[model_key]
        gpus = list(gpus)
        if model_key == "fastpitch1_1":
            self.models_bank[model_key] = FastPitchTrainer(self.logger, self.PROD, gpus, self, websocket=websocket)
        elif model_key == "hifigan":
            self.models_bank[model_key] = HiFiTrainer(self.logger, self.PROD, gpus, self, websocket=websocket)
        else:
            raise KeyError(f"{model_key}: only the fastpitch1_1 and hifigan trainers are part of this build")
        return self.models_bank[model_key]

    def models(self, key):
        return self.models_bank[key]
