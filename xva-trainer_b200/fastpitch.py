"""B200-native FastPitch 1.1 training path: the reference's ``FastPitch`` / ``FastPitchLoss`` / ``Lamb`` surface, with
every tensor operation issued through the C ABI of libxva_b200.so (tcgen05 tap-GEMMs + HBM row kernels).

Mirrors (names, argument meaning, outputs) of the reference, paths relative to python/fastpitch1_1/ :
    FastPitch            fastpitch/model.py:125-390      (forward -> the 13-list of model.py:388-390)
    FastPitchLoss        fastpitch/loss_function.py:29-154
    stage-1 aligner      fastpitch/attention.py:171-220 (ConvAttention), fastpitch/alignment.py:79-118 (MAS),
                         fastpitch/attn_loss_function.py:20-54 (AttentionCTCLoss, AttentionBinarizationLoss)
    Lamb                 lamb.py:8-106                    (+ clip_grad_norm_ of xva_train.py:857)
    noam learning rate   xva_train.py:1252-1261

What differs by design (B200-first):
  * activations are channels-last [B, T, C] fp32 in HBM; the dense contractions run as tf32 tcgen05 MMAs with fp32
    accumulation, the rest as coalesced fp32 row kernels; nothing here calls a torch math op on an activation;
  * there is no autograd graph: forward() records the activations it needs and backward() walks the layers in
    reverse, writing weight gradients straight into a flat gradient arena that the multi-tensor LAMB consumes;
  * parameters live in one flat fp32 arena in the layout the kernels read (conv weights as [taps, Cout, Cin]);
    state_dict() / load_state_dict() convert to and from the reference's keys and shapes, so checkpoints interchange;
  * padding must be trailing (tokens != 0 is a prefix of each row) -- true for every batch the reference's collate
    builds (data_function.py:565-695).
There is no CPU or eager fallback: without the CUDA library every call raises.
"""
import math
import os
from collections import OrderedDict

import torch

from . import capi, ops

D_MODEL, D_HEAD, D_INNER, N_LAYERS, N_MEL, N_SYMBOLS, D_PRED = 384, 64, 1536, 6, 80, 148, 256
P_DROP = 0.1            # p_{in,out}_fft_dropout, dropatt, predictor dropout: model.py:131-179
K3 = (-1, 0, 1)         # row shifts of a kernel-3, padding-1 convolution
_ALIGN = 64             # arena entries start on 256-byte boundaries (TMA needs 16)


def _state_spec():
    """(key, reference shape) for every FastPitch().state_dict() entry, in the reference's order (model.py:125-265)."""
    spec = [("pitch_mean", (1,)), ("pitch_std", (1,))]

    def fft(prefix, embed):
        out = []
        if embed:
            out.append((f"{prefix}.word_emb.weight", (N_SYMBOLS, D_MODEL)))
        out.append((f"{prefix}.pos_emb.inv_freq", (D_MODEL // 2,)))
        for i in range(N_LAYERS):
            p = f"{prefix}.layers.{i}"
            out += [(f"{p}.dec_attn.qkv_net.weight", (3 * D_HEAD, D_MODEL)), (f"{p}.dec_attn.qkv_net.bias", (3 * D_HEAD,)),
                    (f"{p}.dec_attn.o_net.weight", (D_MODEL, D_HEAD)),
                    (f"{p}.dec_attn.layer_norm.weight", (D_MODEL,)), (f"{p}.dec_attn.layer_norm.bias", (D_MODEL,)),
                    (f"{p}.pos_ff.CoreNet.0.weight", (D_INNER, D_MODEL, 3)), (f"{p}.pos_ff.CoreNet.0.bias", (D_INNER,)),
                    (f"{p}.pos_ff.CoreNet.2.weight", (D_MODEL, D_INNER, 3)), (f"{p}.pos_ff.CoreNet.2.bias", (D_MODEL,)),
                    (f"{p}.pos_ff.layer_norm.weight", (D_MODEL,)), (f"{p}.pos_ff.layer_norm.bias", (D_MODEL,))]
        return out

    def predictor(prefix):
        out = []
        for i, cin in enumerate((D_MODEL, D_PRED)):
            p = f"{prefix}.layers.{i}"
            out += [(f"{p}.conv.weight", (D_PRED, cin, 3)), (f"{p}.conv.bias", (D_PRED,)),
                    (f"{p}.norm.weight", (D_PRED,)), (f"{p}.norm.bias", (D_PRED,))]
        out += [(f"{prefix}.fc.weight", (1, D_PRED)), (f"{prefix}.fc.bias", (1,))]
        return out

    spec += fft("encoder", True)
    spec += predictor("duration_predictor")
    spec += fft("decoder", False)
    spec += predictor("pitch_predictor")
    spec += [("pitch_emb.weight", (D_MODEL, 1, 3)), ("pitch_emb.bias", (D_MODEL,))]
    spec += predictor("energy_predictor")
    spec += [("energy_emb.weight", (D_MODEL, 1, 3)), ("energy_emb.bias", (D_MODEL,))]
    spec += [("proj.weight", (N_MEL, D_MODEL)), ("proj.bias", (N_MEL,))]
    spec += [("attention.query_proj.0.conv.weight", (160, 80, 3)), ("attention.query_proj.0.conv.bias", (160,)),
             ("attention.query_proj.2.conv.weight", (80, 160, 1)), ("attention.query_proj.2.conv.bias", (80,)),
             ("attention.query_proj.4.conv.weight", (80, 80, 1)), ("attention.query_proj.4.conv.bias", (80,)),
             ("attention.attn_proj.weight", (1, 80, 1, 1)), ("attention.attn_proj.bias", (1,)),
             ("attention.key_proj.0.conv.weight", (768, 384, 3)), ("attention.key_proj.0.conv.bias", (768,)),
             ("attention.key_proj.2.conv.weight", (80, 768, 1)), ("attention.key_proj.2.conv.bias", (80,))]
    return spec


BUFFERS = ("pitch_mean", "pitch_std", "encoder.pos_emb.inv_freq", "decoder.pos_emb.inv_freq")


def _packed_shape(key, shape):
    """Kernel-side layout of a reference tensor: Conv1d [Cout, Cin, k] -> [k, Cout, Cin]; Linear [N, K] -> [1, N, K];
    the 1 -> C scalar convs and C -> 1 projections are flat."""
    if key.endswith("_emb.weight") and len(shape) == 3 and shape[1] == 1:
        return (shape[0], shape[2])
    if key.endswith(".fc.weight"):
        return (shape[1],)
    if len(shape) == 3:
        return (shape[2], shape[0], shape[1])
    if len(shape) == 2 and not key.endswith("word_emb.weight"):
        return (1, shape[0], shape[1])
    return tuple(shape)


def _to_packed(key, ref):
    shape = tuple(ref.shape)
    ps = _packed_shape(key, shape)
    if len(shape) == 3 and len(ps) == 3 and not key.endswith("_emb.weight"):
        return ref.permute(2, 0, 1).contiguous()
    return ref.reshape(ps)


def _to_ref(key, packed, ref_shape):
    if len(ref_shape) == 3 and packed.dim() == 3 and not key.endswith("_emb.weight"):
        return packed.permute(1, 2, 0).contiguous()
    return packed.reshape(ref_shape).clone()


def trainable_keys(stage):
    """Parameter keys that receive gradients in a training stage (freezing of xva_train.py:607-669)."""
    keys = [k for k, _ in _state_spec() if k not in BUFFERS]

    def under(*prefixes):
        return [k for k in keys if any(k.startswith(p + ".") for p in prefixes)]

    if stage == 1:
        # The trainer un-freezes encoder + attention (xva_train.py:607-620), but the stage-1 loss only reaches the token
        # embedding (the keys are encoder.word_emb(inputs), model.py:299; enc_out is not used) and the two projection
        # stacks: every other tensor keeps grad = None in the reference, and Lamb.step skips those (lamb.py:52-53) --
        # no moment update, no weight decay. attention.attn_proj is never called by ConvAttention.forward.
        return ["encoder.word_emb.weight"] + [k for k in under("attention") if ".attn_proj." not in k]
    if stage == 2:
        return under("encoder", "duration_predictor")
    if stage == 3:
        return [k for k in keys if not (k.startswith("attention.") or k.startswith("duration_predictor."))]
    if stage == 4:
        return under("encoder", "decoder", "energy_emb", "proj")
    raise ValueError(f"training stage {stage}: FastPitch has stages 1-4")


class _Arena:
    """Flat fp32 storage for parameters (p), gradients (g) and the two LAMB moments (m, v), one layout for all four."""

    def __init__(self, device):
        self.spec = [(k, s) for k, s in _state_spec() if k not in BUFFERS]
        self.offset, self.pshape, self.rshape = {}, {}, {}
        off = 0
        for k, s in self.spec:
            ps = _packed_shape(k, s)
            n = int(math.prod(ps))
            self.offset[k], self.pshape[k], self.rshape[k] = off, ps, tuple(s)
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.p = torch.zeros(off, device=device, dtype=torch.float32)   # fp32 master parameters (what LAMB updates)
        self.w = torch.zeros(off, device=device, dtype=torch.float32)   # the same, rounded to tf32: GEMM B operands
        self.g = torch.zeros(off, device=device, dtype=torch.float32)
        self.m = None
        self.v = None

    def view(self, flat, key):
        o, ps = self.offset[key], self.pshape[key]
        return flat[o:o + int(math.prod(ps))].view(ps)


class _NS:
    """attribute bag"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class FastPitch(torch.nn.Module):
    """Drop-in for the reference ``FastPitch`` (fastpitch/model.py:125): same constructor argument, ``training_stage``
    attribute, ``pitch_mean`` / ``pitch_std`` buffers, ``forward(inputs_x, ...)`` -> 13-list, state_dict keys and shapes.

    Training adds ``backward(criterion, scale)`` (the hand-written reverse pass; there is no autograd graph)."""

    def __init__(self, logger=None, device=None, seed=1234):
        super().__init__()
        self.logger = logger
        self.training_stage = 3
        self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device_.type != "cuda":
            raise capi.XvaError("FastPitch (B200 build) needs a CUDA device: there is no CPU path")
        capi.load()
        capi.call("xva_device_check", self.device_.index or 0)
        self.arena = _Arena(self.device_)
        dev = self.device_
        self.pitch_mean = torch.zeros(1, device=dev)
        self.pitch_std = torch.zeros(1, device=dev)
        self.inv_freq = (1.0 / (10000 ** (torch.arange(0.0, D_MODEL, 2.0) / D_MODEL))).to(dev)
        self.p_drop = P_DROP
        self.fuse_ln = os.environ.get("XVA_FUSE_LN", "0") == "1"   # LayerNorm inside the GEMM epilogue (measured slower)
        # softmax backward inside the dP GEMM epilogue: parity-green but measured slower (13.84 vs 13.59 ms/step: the
        # epilogue-bound GEMM gets heavier than the HBM-speed row kernel it replaces), so off by default
        self.fuse_softmax_bwd = os.environ.get("XVA_FUSE_SOFTMAX_BWD", "0") == "1"
        # one fused tcgen05 kernel per direction for scores / softmax / dropout / P.V (XVA_FUSED_ATTN=0: the unfused chain of
        # six K = 64 GEMMs + softmax kernels with the [B, T, T] tensors in HBM, kept as the A/B and parity partner)
        self.fused_attn = os.environ.get("XVA_FUSED_ATTN", "1") != "0"
        # EXPERIMENT, off by default, not yet measured (round-2 item): issue every FFT-block weight
        # gradient on a second stream. It only reads activations the forward saved and the layer's output gradient and
        # accumulates into the gradient arena, so it is independent of the input-gradient GEMM that follows it on the main
        # stream; the persistent CTAs of one kernel that run out of tiles (224 row tiles on 148 SMs = 1.51 waves) free
        # their SMs for the other kernel instead of idling until the launch ends.
        self.bwd_streams = None      # None: by XVA_BWD_STREAMS (default auto = inside CUDA-graph capture); bool: forced
        self._side = None
        self._side_used = False
        self.seed = int(seed)
        self.step_counter = torch.zeros(1, device=dev, dtype=torch.int64)  # device-side dropout counter (uint64 bits)
        self._site = 0
        self._ctx = None
        self._bind()
        self.reset_parameters(seed)

    # ------------------------------------------------------------------------------------------ parameters
    def _bind(self):
        """Name the packed parameter / gradient views the kernels read."""
        A = self.arena
        G = lambda k: A.view(A.g, k)

        def W(k):
            """Weight matrices that are tensor-core operands are read from the tf32-rounded copy (the MMA would
            otherwise truncate them); biases, LayerNorm parameters and the small non-GEMM weights stay fp32."""
            is_gemm_operand = k.endswith(".weight") and len(A.pshape[k]) == 3
            return A.view(A.w if is_gemm_operand else A.p, k)

        def fft(prefix):
            layers = []
            for i in range(N_LAYERS):
                p = f"{prefix}.layers.{i}"
                names = dict(qkv_w=f"{p}.dec_attn.qkv_net.weight", qkv_b=f"{p}.dec_attn.qkv_net.bias",
                             o_w=f"{p}.dec_attn.o_net.weight", ln1_g=f"{p}.dec_attn.layer_norm.weight",
                             ln1_b=f"{p}.dec_attn.layer_norm.bias", w1=f"{p}.pos_ff.CoreNet.0.weight",
                             b1=f"{p}.pos_ff.CoreNet.0.bias", w2=f"{p}.pos_ff.CoreNet.2.weight",
                             b2=f"{p}.pos_ff.CoreNet.2.bias", ln2_g=f"{p}.pos_ff.layer_norm.weight",
                             ln2_b=f"{p}.pos_ff.layer_norm.bias")
                layers.append(_NS(w=_NS(**{n: W(k) for n, k in names.items()}), g=_NS(**{n: G(k) for n, k in names.items()})))
            return layers

        def predictor(prefix):
            names = {}
            for i in range(2):
                p = f"{prefix}.layers.{i}"
                names.update({f"w{i}": f"{p}.conv.weight", f"b{i}": f"{p}.conv.bias", f"g{i}": f"{p}.norm.weight",
                              f"be{i}": f"{p}.norm.bias"})
            names.update(fc_w=f"{prefix}.fc.weight", fc_b=f"{prefix}.fc.bias")
            return _NS(w=_NS(**{n: W(k) for n, k in names.items()}), g=_NS(**{n: G(k) for n, k in names.items()}))

        self.enc_layers, self.dec_layers = fft("encoder"), fft("decoder")
        self.pred = {n: predictor(f"{n}_predictor") for n in ("duration", "pitch", "energy")}
        misc = dict(emb="encoder.word_emb.weight", pitch_emb_w="pitch_emb.weight", pitch_emb_b="pitch_emb.bias",
                    energy_emb_w="energy_emb.weight", energy_emb_b="energy_emb.bias", proj_w="proj.weight",
                    proj_b="proj.bias")
        self.w = _NS(**{n: W(k) for n, k in misc.items()})
        self.g = _NS(**{n: G(k) for n, k in misc.items()})
        att = {}
        for stack, idx in (("query_proj", (0, 2, 4)), ("key_proj", (0, 2))):
            for i in idx:
                att[f"{stack[0]}{i}_w"] = f"attention.{stack}.{i}.conv.weight"
                att[f"{stack[0]}{i}_b"] = f"attention.{stack}.{i}.conv.bias"
        self.att = _NS(w=_NS(**{n: W(k) for n, k in att.items()}), g=_NS(**{n: G(k) for n, k in att.items()}))

    def reset_parameters(self, seed=1234):
        """Seeded init with torch's default fan-in scales (what nn.Conv1d / nn.Linear / nn.Embedding / nn.LayerNorm do
        in the reference constructor); generated on the CPU so it is identical on every rank and machine."""
        g = torch.Generator().manual_seed(int(seed))
        sd = OrderedDict()
        for key, shape in _state_spec():
            if key in BUFFERS:
                continue
            if ".layer_norm." in key or ".norm." in key:
                sd[key] = torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
            elif key.endswith("word_emb.weight"):
                w = torch.randn(shape, generator=g)
                w[0] = 0.0
                sd[key] = w
            else:
                wshape = shape
                if key.endswith(".bias"):
                    wkey = key[:-4] + "weight"
                    wshape = dict(_state_spec())[wkey]
                fan_in = int(math.prod(wshape[1:]))
                bound = 1.0 / math.sqrt(fan_in)
                sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        self.load_state_dict(sd, strict=False)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """The reference's 185 keys with the reference's shapes (conv weights back in [Cout, Cin, k])."""
        out = OrderedDict() if destination is None else destination
        A = self.arena
        for key, shape in _state_spec():
            if key == "pitch_mean":
                t = self.pitch_mean.clone()
            elif key == "pitch_std":
                t = self.pitch_std.clone()
            elif key.endswith("inv_freq"):
                t = self.inv_freq.clone()
            else:
                t = _to_ref(key, A.view(A.p, key), tuple(shape))
            out[prefix + key] = t
        return out

    def load_state_dict(self, state_dict, strict=True):
        A = self.arena
        known = dict(_state_spec())
        missing = [k for k in known if k not in state_dict]
        unexpected = [k for k in state_dict if k not in known]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        with torch.no_grad():
            for key, t in state_dict.items():
                if key not in known:
                    continue
                t = t.detach().to(torch.float32)
                if tuple(t.shape) != tuple(known[key]):
                    raise RuntimeError(f"load_state_dict: {key} has shape {tuple(t.shape)}, expected {tuple(known[key])}")
                if key == "pitch_mean":
                    self.pitch_mean.copy_(t)
                elif key == "pitch_std":
                    self.pitch_std.copy_(t)
                elif key.endswith("inv_freq"):
                    self.inv_freq.copy_(t)
                else:
                    A.view(A.p, key).copy_(_to_packed(key, t))
            ops.round_tf32_(A.p, A.w)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def grads(self, keys=None):
        """{reference key: gradient in the reference's shape} for the given (default: all) parameter keys."""
        A = self.arena
        keys = [k for k, _ in A.spec] if keys is None else keys
        return {k: _to_ref(k, A.view(A.g, k), A.rshape[k]) for k in keys}

    def zero_grad(self, set_to_none=True):
        self.arena.g.zero_()

    def parameters(self, recurse=True):  # the arena is the one parameter tensor (what an optimizer / DDP bucket sees)
        return iter([self.arena.p])

    def to(self, *a, **k):
        return self

    # ------------------------------------------------------------------------------------------ dropout bookkeeping
    def _drop(self):
        """(p, seed) for the next dropout site of this pass; p = 0 outside training. Sites are numbered in call order,
        so backward re-derives the same masks from the numbers it saved."""
        self._site += 1
        p = self.p_drop if self.training else 0.0
        return p, (self.seed * 0x9E3779B1 + self._site * 0x85EBCA77) & 0xFFFFFFFFFFFF

    # ------------------------------------------------------------------------------------------ FFT block
    def _layer_fwd(self, x, lens, L, save):
        """TransformerLayer.forward, transformer.py:164-171 = MultiHeadAttn :100-152 + PositionwiseConvFF :59-77."""
        B, T, _ = x.shape
        sd = self.step_counter
        qkv = ops.conv_fwd(x, L.w.qkv_w, bias=L.w.qkv_b, round_out=True)
        q, k, v = qkv[..., :D_HEAD], qkv[..., D_HEAD:2 * D_HEAD], qkv[..., 2 * D_HEAD:]
        Tp = (T + 31) // 32 * 32
        p_att, seed_att = self._drop()
        P = Pd = lse = None
        if self.fused_attn:
            # scores, key mask, softmax, dropout and P.V in one kernel, S / P in tensor memory (csrc/attn_fused.cu); only
            # the row log-sum-exp is kept for the backward, which recomputes P
            vec, lse = ops.attn_fwd(qkv, lens, 1.0 / math.sqrt(D_HEAD), p_att, seed_att, sd, Tp)
        else:
            s = torch.empty(B, T, Tp, device=x.device, dtype=torch.float32)
            ops.bmm_nt(q, k, alpha=1.0 / math.sqrt(D_HEAD), out=s[..., :T])
            P, Pd = ops.softmax_fwd(s, lens, T, p_att, seed_att, sd)
            del s
            vec = ops.bmm_nn(Pd[..., :T], v, round_out=True)
        p1, seed1 = self._drop()
        # The post-LN of both sub-blocks is a separate HBM pass (xva_layernorm_fwd) over the pre-LN sum the GEMM epilogue
        # wrote (bias + dropout + residual): a 384-column LayerNorm epilogue needs the whole row in one accumulator, which
        # cannot be double-buffered in TMEM and left 54 % (T = 160) or 27 % (T = 880) of the SMs without a tile.
        if self.fuse_ln:
            y1, sv1 = ops.conv_fwd(vec, L.w.o_w, residual=x, ln=(L.w.ln1_g, L.w.ln1_b), save_ln=True, lens=lens,
                                   drop_p=p1, seed=seed1, seed_dev=sd, round_out=True)
        else:
            pre1 = ops.conv_fwd(vec, L.w.o_w, residual=x, drop_p=p1, seed=seed1, seed_dev=sd)
            y1, sv1 = ops.layernorm_fwd(pre1, L.w.ln1_g, L.w.ln1_b, lens)
        h = ops.conv_fwd(y1, L.w.w1, K3, bias=L.w.b1, relu=True, round_out=True)
        p2, seed2 = self._drop()
        if self.fuse_ln:
            y2, sv2 = ops.conv_fwd(h, L.w.w2, K3, bias=L.w.b2, residual=y1, ln=(L.w.ln2_g, L.w.ln2_b), save_ln=True,
                                   lens=lens, drop_p=p2, seed=seed2, seed_dev=sd, round_out=True)
        else:
            pre2 = ops.conv_fwd(h, L.w.w2, K3, bias=L.w.b2, residual=y1, drop_p=p2, seed=seed2, seed_dev=sd)
            y2, sv2 = ops.layernorm_fwd(pre2, L.w.ln2_g, L.w.ln2_b, lens)
        if save is not None:
            save.append(_NS(x=x, qkv=qkv, P=P, Pd=Pd, lse=lse, vec=vec, sv1=sv1, y1=y1, h=h, sv2=sv2, T=T, att=(p_att, seed_att),
                            d1=(p1, seed1), d2=(p2, seed2)))
        return y2

    def _wgrad_side(self, fn, *inputs):
        """Run ``fn`` (weight / bias gradient launches) on the side stream when the two-stream backward is on. ``inputs``
        are record_stream()-ed: the caching allocator does not hand their memory to a main-stream kernel before the side
        stream is done with it (inside a graph capture such blocks are simply not reused). The main stream does not wait
        for the side stream until the gradients are consumed (_join_side)."""
        if not self._side_on():
            return fn()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device_)
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            fn()
        for t in inputs:
            t.record_stream(self._side)
        self._side_used = True

    def _side_on(self):
        """Two-stream backward: measured 14.07 -> 13.69 ms / step inside the replayed graph and 14.86 -> 14.34 launched
        eagerly at 32 x 880 x 160 (profiles/r02_streams_ab.txt)."""
        if self.bwd_streams is not None:
            return bool(self.bwd_streams)
        v = os.environ.get("XVA_BWD_STREAMS", "auto").strip().lower()
        if v in ("", "auto"):
            return ops.capturing()
        return v not in ("0", "false", "off")

    def _pred_side(self):
        """Stream of the pitch / energy predictors when they run next to the decoder (forward and backward)."""
        if getattr(self, "_pred_stream", None) is None:
            self._pred_stream = torch.cuda.Stream(device=self.device_)
        return self._pred_stream

    def _join_side(self):
        """Before anything reads the gradient arena from the main stream: GradSync.ready, the end of backward()."""
        if self._side is not None and self._side_used:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_used = False

    def _layer_bwd(self, dy, lens, L, c, need_dx=True):
        sd = self.step_counter
        x, qkv = c.x, c.qkv
        B, T, _ = x.shape
        q, k, v = qkv[..., :D_HEAD], qkv[..., D_HEAD:2 * D_HEAD], qkv[..., 2 * D_HEAD:]
        # ---- PositionwiseConvFF
        dx2, dbr2 = ops.layernorm_bwd(dy, c.sv2, L.w.ln2_g, lens, L.g.ln2_g, L.g.ln2_b, dbias=L.g.b2, want_drop=True,
                                      drop_pre_p=c.d2[0], seed_pre=c.d2[1], seed_dev=sd)
        self._wgrad_side(lambda: ops.conv_wgrad(dbr2, c.h, K3, out=L.g.w2, accumulate=True), dbr2)
        dh = ops.conv_dgrad(dbr2, L.w.w2, K3, gate=c.h, round_out=True)
        del dbr2

        def w1_grads():
            ops.conv_wgrad(dh, c.y1, K3, out=L.g.w1, accumulate=True)
            ops.colsum_(B * T, D_INNER, D_INNER, dh, L.g.b1)

        self._wgrad_side(w1_grads, dh)
        dy1 = ops.conv_dgrad(dh, L.w.w1, K3, residual=dx2)
        del dh, dx2
        # ---- MultiHeadAttn
        dx1, dbr1 = ops.layernorm_bwd(dy1, c.sv1, L.w.ln1_g, lens, L.g.ln1_g, L.g.ln1_b, dbias=None, want_drop=True,
                                      drop_pre_p=c.d1[0], seed_pre=c.d1[1], seed_dev=sd)
        self._wgrad_side(lambda: ops.conv_wgrad(dbr1, c.vec, (0,), out=L.g.o_w, accumulate=True), dbr1)
        dvec = ops.conv_dgrad(dbr1, L.w.o_w, round_out=True)
        dqkv = torch.empty_like(qkv)
        Tp = (T + 31) // 32 * 32
        if c.lse is not None:
            ops.attn_bwd(qkv, dvec, c.vec, c.lse, lens, 1.0 / math.sqrt(D_HEAD), c.att[0], c.att[1], sd, Tp, out=dqkv)
            dP = None
        elif self.fuse_softmax_bwd:
            dP = torch.empty(B, T, Tp, device=x.device, dtype=torch.float32)
            # softmax + attention-dropout backward inside the epilogue of dP = dO.V^T: the row term sum_j P_d dP_d
            # equals dO.O (O = P_d V = vec), so dS is element-wise in the tile and the score-sized tensor is written
            # once instead of written, read twice and written again.
            if Tp > T:
                dP[..., T:].zero_()
            D = ops.rowdot2(dvec, c.vec)
            ops.bmm_nt(dvec, v, alpha=1.0 / math.sqrt(D_HEAD), out=dP[..., :T], round_out=True,
                       softmax_bwd=(c.P, D, c.att[0], c.att[1], sd))
            ops.bmm_tn(c.Pd[..., :T], dvec, out=dqkv[..., 2 * D_HEAD:], round_out=True)
        else:
            dP = torch.empty(B, T, Tp, device=x.device, dtype=torch.float32)
            ops.bmm_nt(dvec, v, out=dP[..., :T])
            ops.bmm_tn(c.Pd[..., :T], dvec, out=dqkv[..., 2 * D_HEAD:], round_out=True)
            ops.softmax_bwd_(c.P, dP, T, 1.0 / math.sqrt(D_HEAD), c.att[0], c.att[1], sd)
        if dP is not None:
            ops.bmm_nn(dP[..., :T], k, out=dqkv[..., :D_HEAD], round_out=True)
            ops.bmm_tn(dP[..., :T], q, out=dqkv[..., D_HEAD:2 * D_HEAD], round_out=True)
        del dP
        def qkv_grads():
            ops.conv_wgrad(dqkv, x, (0,), out=L.g.qkv_w, accumulate=True)
            ops.colsum_(B * T, 3 * D_HEAD, 3 * D_HEAD, dqkv, L.g.qkv_b)

        self._wgrad_side(qkv_grads, dqkv)
        return ops.conv_dgrad(dqkv, L.w.qkv_w, residual=dx1) if need_dx else None

    # ------------------------------------------------------------------------------------------ temporal predictor
    def _pred_fwd(self, x, lens, P, save):
        """TemporalPredictor.forward, model.py:118-122 (ConvReLUNorm common/layers.py:94-97). x is already zero on
        padded rows (= enc_out * mask). -> [B, Tt]"""
        pa, sa = self._drop()
        h1, s1 = ops.conv_fwd(x, P.w.w0, K3, bias=P.w.b0, relu=True, ln=(P.w.g0, P.w.be0), save_ln=True, drop_p=pa,
                              drop_post=True, seed=sa, seed_dev=self.step_counter, round_out=True)
        pb, sb = self._drop()
        h2, s2 = ops.conv_fwd(h1, P.w.w1, K3, bias=P.w.b1, relu=True, ln=(P.w.g1, P.w.be1), save_ln=True, drop_p=pb,
                              drop_post=True, seed=sb, seed_dev=self.step_counter)
        out = ops.rowdot_fwd(h2, P.w.fc_w, P.w.fc_b, lens)
        if save is not None:
            save.update(x=x, h1=h1, s1=s1, h2=h2, s2=s2, da=(pa, sa), db=(pb, sb))
        return out

    def _pred_bwd(self, dpred, lens, P, c, residual=None):
        """-> gradient wrt the predictor input (+ residual), zero on padded rows."""
        sd = self.step_counter
        dh2 = ops.rowdot_bwd(dpred, c["h2"], P.w.fc_w, lens, P.g.fc_w, P.g.fc_b)
        dc2, _ = ops.layernorm_bwd(dh2, c["s2"], P.w.g1, None, P.g.g1, P.g.be1, dbias=P.g.b1, drop_post_p=c["db"][0],
                                   seed_post=c["db"][1], seed_dev=sd, relu_gate=True)
        ops.conv_wgrad(dc2, c["h1"], K3, out=P.g.w1, accumulate=True)
        dh1 = ops.conv_dgrad(dc2, P.w.w1, K3)
        dc1, _ = ops.layernorm_bwd(dh1, c["s1"], P.w.g0, None, P.g.g0, P.g.be0, dbias=P.g.b0, drop_post_p=c["da"][0],
                                   seed_post=c["da"][1], seed_dev=sd, relu_gate=True)
        ops.conv_wgrad(dc1, c["x"], K3, out=P.g.w0, accumulate=True)
        return ops.conv_dgrad(dc1, P.w.w0, K3, residual=residual, lens=lens)

    # ------------------------------------------------------------------------------------------ stage-1 alignment
    def binarize_attention_parallel(self, attn, in_lens, out_lens):
        """FastPitch.binarize_attention_parallel, model.py:283-294: hard (Viterbi) alignment of the soft attention
        attn [B, 1, max_mel_len, max_text_len]. The reference copies attn to the host, runs numba's b_mas and copies the
        result back; here it is one kernel on the device (xva_mas_width1), bit-identical path on the same fp32 values."""
        with torch.no_grad():
            hard, _ = ops.mas_width1(attn, in_lens, out_lens)
        return hard

    binarize_attention = binarize_attention_parallel          # model.py:267-281: the same result, item by item

    # ------------------------------------------------------------------------------------------ forward
    def _forward_stage1(self, inputs_x):
        """Training stage 1: FastPitch.get_alignment_durations, model.py:298-323, and the return of model.py:356-360.
        ConvAttention (attention.py:171-220): the two projection stacks on the tap-GEMM (channels-last, the k=3 halo and
        the utterance boundaries are TMA zero fill), the Gaussian score + log_softmax + prior + masked softmax in one
        kernel (xva_attn_score_fwd), the Viterbi alignment on the device (xva_mas_width1; the reference copies the
        attention to the host for numba and back). The reference also runs the encoder stack here (model.py:345) but
        returns nothing that depends on it in stage 1, so it is skipped."""
        (inputs, input_lens, mel_tgt, mel_lens, _, _, _, attn_prior, _, max_inp_lengths, _, _) = inputs_x
        if attn_prior is None:
            raise ValueError("training stage 1 needs the alignment prior (inputs_x[7], data_function.py:88-99)")
        dev = self.device_
        B, Tt = inputs.shape
        Tm = mel_tgt.shape[2]
        tokens = inputs.to(torch.int64).contiguous()
        in_lens32 = input_lens.to(torch.int32).contiguous()
        out_lens32 = mel_lens.to(torch.int32).contiguous()
        A = self.att.w
        # 80-channel tensors live in rows of 96 floats with a zero tail: they are MN-major operands (read in 32-column
        # chunks) of the weight-gradient GEMMs of the backward pass
        wide = lambda rows: torch.zeros(B, rows, 96, device=dev, dtype=torch.float32)
        # keys: key_proj(encoder.word_emb(inputs)) -- Conv1d(384, 768, 3) -> ReLU -> Conv1d(768, 80, 1), attention.py:95-107
        text_emb = ops.embed_pos(tokens, self.w.emb, None, None, None, B, Tt, D_MODEL)
        kh = ops.conv_fwd(text_emb, A.k0_w, K3, bias=A.k0_b, relu=True, round_out=True)
        kbuf = wide(Tt)
        k_enc = ops.conv_fwd(kh, A.k2_w, bias=A.k2_b, out=kbuf[..., :N_MEL])
        # queries: query_proj(mel) -- Conv1d(80, 160, 3) -> ReLU -> Conv1d(160, 80, 1) -> ReLU -> Conv1d(80, 80, 1), :113-128
        melp = wide(Tm)
        mel = melp[..., :N_MEL]
        mel.copy_(mel_tgt.to(torch.float32).transpose(1, 2))
        ops.round_tf32_(melp.view(-1), melp.view(-1))
        qh1 = ops.conv_fwd(mel, A.q0_w, K3, bias=A.q0_b, relu=True, round_out=True)
        qh2 = ops.conv_fwd(qh1, A.q2_w, bias=A.q2_b, relu=True, round_out=True, out=wide(Tm)[..., :N_MEL])
        q_enc = ops.conv_fwd(qh2, A.q4_w, bias=A.q4_b, out=wide(Tm)[..., :N_MEL])
        # score, log_softmax + prior, masked softmax (attention.py:203-219)
        logprob, soft, prior = ops.attn_score_fwd(q_enc, k_enc, attn_prior, in_lens32)
        attn_hard, durs = ops.mas_width1(soft, in_lens32, out_lens32)
        if self.training:
            self._ctx = _NS(stage=1, B=B, Tt=Tt, Tm=Tm, tokens=tokens, in_lens=in_lens32, text_emb=text_emb, kh=kh, k=k_enc,
                            mel=mel, qh1=qh1, qh2=qh2, q=q_enc, logprob=logprob, prior=prior)
        else:
            self._ctx = None
        shape4 = (B, 1, Tm, Tt)
        return [None, None, None, None, None, None, None, None, soft.view(shape4), attn_hard.view(shape4),
                durs.to(torch.float32), logprob.view(shape4), input_lens]

    def _backward_stage1(self, criterion, scale, kl, ready):
        """Reverse of _forward_stage1 for loss = attn_loss_scale * ctc + kl_weight * binarization (xva_train.py:790-798)."""
        ctx = self._ctx
        A, G = self.att.w, self.att.g
        B, Tt, Tm = ctx.B, ctx.Tt, ctx.Tm
        gctc, a = criterion.grad_seeds(scale)["attn_logprob"]
        if kl is not None and kl[1]:
            g = ops.attn_grad_combine(gctc, a, *kl[0].saved(), bw=scale * float(kl[1]))
        else:
            g = ops.attn_grad_combine(gctc, a)
        dq, dk = ops.attn_score_bwd(g, ctx.logprob, ctx.prior, ctx.q, ctx.k)
        del g
        # query_proj, last layer first. The 80 -> 80 layer's weight is copied into rows of 96 floats: the input-gradient
        # GEMM reads it MN-major in 32-column chunks.
        ops.conv_wgrad(dq, ctx.qh2, (0,), out=G.q4_w, accumulate=True)
        ops.colsum_(B * Tm, N_MEL, dq.stride(1), dq, G.q4_b)
        w4 = torch.zeros(1, N_MEL, 96, device=dq.device, dtype=torch.float32)
        w4[..., :N_MEL].copy_(A.q4_w)
        dh2 = ops.conv_dgrad(dq, w4[..., :N_MEL], gate=ctx.qh2, round_out=True,
                             out=torch.zeros(B, Tm, 96, device=dq.device, dtype=torch.float32)[..., :N_MEL])
        ops.conv_wgrad(dh2, ctx.qh1, (0,), out=G.q2_w, accumulate=True)
        ops.colsum_(B * Tm, N_MEL, dh2.stride(1), dh2, G.q2_b)
        dh1 = ops.conv_dgrad(dh2, A.q2_w, gate=ctx.qh1, round_out=True)
        ops.conv_wgrad(dh1, ctx.mel, K3, out=G.q0_w, accumulate=True)
        ops.colsum_(B * Tm, dh1.shape[2], dh1.shape[2], dh1, G.q0_b)
        del dq, dh2, dh1
        # key_proj and the token embedding
        ops.conv_wgrad(dk, ctx.kh, (0,), out=G.k2_w, accumulate=True)
        ops.colsum_(B * Tt, N_MEL, dk.stride(1), dk, G.k2_b)
        dkh = ops.conv_dgrad(dk, A.k2_w, gate=ctx.kh, round_out=True)
        ops.conv_wgrad(dkh, ctx.text_emb, K3, out=G.k0_w, accumulate=True)
        ops.colsum_(B * Tt, dkh.shape[2], dkh.shape[2], dkh, G.k0_b)
        demb = ops.conv_dgrad(dkh, A.k0_w, K3)
        ops.embed_bwd_(ctx.tokens, demb, self.g.emb)
        ready("attention", flush=True)              # two slices at opposite ends of the arena: sent separately
        ready("encoder.word_emb", flush=True)
        self._ctx = None

    def forward(self, inputs_x, use_gt_pitch=True, use_dur_tgt=False, pace=1.0, max_duration=75, host_lens=None):
        """FastPitch.forward, model.py:325-390, training stages 1-4. ``inputs_x`` is the 12-list of
        data_function.py:737-738. ``host_lens`` = (mel_max_len, max(dec_lens)) as Python ints lets a caller that already
        knows them on the host (the collate does) skip the two device->host reads the reference makes (model.py:330,
        :75). Returns the reference's 13-list; with the module in training mode the activations backward() needs are
        kept until the next forward()."""
        (inputs, input_lens, mel_tgt, mel_lens, pitch_dense, energy_dense, speaker, attn_prior, durs_padded,
         max_inp_lengths, max_mel_lengths, audiopaths) = inputs_x
        stage = int(self.training_stage)
        if stage not in (1, 2, 3, 4):
            trainable_keys(stage)
        self._site = 0
        if stage == 1:
            return self._forward_stage1(inputs_x)
        if not use_gt_pitch:
            raise NotImplementedError("use_gt_pitch=False is the inference path (FastPitch.infer), not part of training")
        dev = self.device_
        self._site = 0
        ctx = _NS(stage=stage, enc=[], dec=[], preds={}) if self.training else None
        save = (lambda name: ctx.preds.setdefault(name, {})) if ctx is not None else (lambda name: None)
        B, Tt = inputs.shape
        tokens = inputs.to(torch.int64).contiguous()
        in_lens32 = input_lens.to(torch.int32)

        # ---- encoder (FFTransformer.forward, transformer.py:212-243)
        with ops.nvtx("fastpitch.fwd.encoder"):
            x = ops.embed_pos(tokens, self.w.emb, None, None, self.inv_freq, B, Tt, D_MODEL)
            for L in self.enc_layers:
                x = self._layer_fwd(x, in_lens32, L, ctx.enc if ctx is not None else None)
        enc_out = x
        dur_tgt = durs_padded
        if ctx is not None:
            ctx.tokens, ctx.in_lens, ctx.B, ctx.Tt = tokens, in_lens32, B, Tt

        if stage == 2:
            log_dur_pred = self._pred_fwd(enc_out, in_lens32, self.pred["duration"], save("duration"))
            dur_pred = torch.clamp(torch.exp(log_dur_pred) - 1, 0, max_duration)  # returned for logging only
            self._ctx = ctx
            return [None, None, dur_pred, log_dur_pred, None, None, None, None, None, None, dur_tgt, None, input_lens]

        # ---- get_pitch_energy, model.py:394-423
        durs = dur_tgt.to(torch.float32).contiguous()
        # The two predictors read the encoder output and the TARGET pitch / energy only, and nothing but the loss reads
        # their outputs: with the side streams on (inside a captured graph) they run next to the decoder instead of in
        # front of it -- 12 launches over 160-frame rows, 0.35 ms during which 40 of 148 SMs had work
        # (profiles/r02_timeline_fastpitch.txt). Only the stream changes: program order, launches and arguments are the
        # same with the streams off (tests/test_launch_sequence.py).
        pred_par = self._side_on() and ctx is not None and stage == 3 and os.environ.get("XVA_PRED_STREAM", "1") != "0"
        ps = self._pred_side() if pred_par else None

        def on_pred_stream(fn):
            if not pred_par:
                return fn()
            ps.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(ps):
                return fn()

        pitch_pred = on_pred_stream(lambda: self._pred_fwd(enc_out, in_lens32, self.pred["pitch"], save("pitch")).view(B, 1, Tt))
        pitch_tgt = ops.average_pitch(pitch_dense, durs)                                    # [B,1,Tt]
        enc2 = enc_out.clone()
        ops.scalar_conv_add_(enc2, pitch_tgt, self.w.pitch_emb_w, self.w.pitch_emb_b, in_lens32)
        energy_pred = on_pred_stream(lambda: self._pred_fwd(enc2, in_lens32, self.pred["energy"], save("energy")))
        energy_tgt = ops.average_pitch(energy_dense.view(B, 1, -1), durs, log1p=True)       # [B,1,Tt]
        enc3 = enc2.clone()
        ops.scalar_conv_add_(enc3, energy_tgt, self.w.energy_emb_w, self.w.energy_emb_b, in_lens32)
        energy_tgt = energy_tgt.view(B, Tt)

        # ---- regulate_len, model.py:59-79
        if host_lens is None:
            mel_max_len = int(max_mel_lengths[0].item())
        else:
            mel_max_len = int(host_lens[0])
        cum, dec_lens = ops.duration_scan(durs, pace, mel_max_len)
        T_out = int(dec_lens.max().item()) if host_lens is None else min(int(host_lens[1]), mel_max_len)
        regulated = ops.regulate_gather(enc3, cum, T_out)

        # ---- decoder + projection
        with ops.nvtx("fastpitch.fwd.decoder"):
            y = ops.embed_pos(None, None, regulated, dec_lens, self.inv_freq, B, T_out, D_MODEL)
            for L in self.dec_layers:
                y = self._layer_fwd(y, dec_lens, L, ctx.dec if ctx is not None else None)
            mel_out = ops.conv_fwd(y, self.w.proj_w, bias=self.w.proj_b)
        if ctx is not None:
            ctx.dec_out, ctx.cum, ctx.dec_lens, ctx.T_out = y, cum, dec_lens, T_out
            ctx.pitch_tgt, ctx.energy_tgt = pitch_tgt, energy_tgt
        if pred_par:       # the criterion reads the predictions on the launching stream
            cur = torch.cuda.current_stream()
            cur.wait_stream(ps)
            pitch_pred.record_stream(cur)
            energy_pred.record_stream(cur)
        self._ctx = ctx
        dec_mask = (torch.arange(T_out, device=dev)[None, :] < dec_lens[:, None]).unsqueeze(2)
        return [mel_out, dec_mask, None, None, pitch_pred, pitch_tgt, energy_pred, energy_tgt, None, None, dur_tgt, None,
                input_lens]

    # ------------------------------------------------------------------------------------------ inference
    def infer(self, inputs, pace=1.0, dur_tgt=None, pitch_tgt=None, energy_tgt=None, pitch_transform=None, max_duration=75,
              speaker=0):
        """FastPitch.infer, model.py:426-482: free-running synthesis (what the UI's preview / export call). Same
        arguments and 5-tuple: (mel_out [B, 80, T], dec_lens, dur_pred [B, Tt], pitch_pred [B, 1, Tt], energy_pred
        [B, Tt]). There is no speaker embedding in this model (speaker_emb is None, model.py:188-199), so ``speaker`` is
        ignored as in the reference. With ``energy_tgt`` the reference fails on an unbound ``energy_pred`` (:462-467,482);
        here it is returned as None. Runs the forward kernels only; call ``eval()`` first, as the reference's callers do."""
        B, Tt = inputs.shape
        tokens = inputs.to(torch.int64).contiguous()
        in_lens32 = (tokens != 0).sum(1).to(torch.int32)          # enc_mask = (tokens != 0), transformer.py:216-217
        self._site = 0
        x = ops.embed_pos(tokens, self.w.emb, None, None, self.inv_freq, B, Tt, D_MODEL)
        for L in self.enc_layers:
            x = self._layer_fwd(x, in_lens32, L, None)
        enc_out = x
        log_dur_pred = self._pred_fwd(enc_out, in_lens32, self.pred["duration"], None)
        dur_pred = torch.clamp(torch.exp(log_dur_pred) - 1, 0, max_duration)
        pitch_pred = self._pred_fwd(enc_out, in_lens32, self.pred["pitch"], None).view(B, 1, Tt)
        if pitch_transform is not None:
            if float(self.pitch_std[0]) == 0.0:
                mean, std = 218.14, 67.24                         # model.py:447-449 (LJSpeech-1.1 defaults)
            else:
                mean, std = self.pitch_mean[0], self.pitch_std[0]
            pitch_pred = pitch_transform(pitch_pred, in_lens32.to(torch.int64), mean, std)
        src = pitch_pred if pitch_tgt is None else pitch_tgt
        enc2 = enc_out.clone()
        ops.scalar_conv_add_(enc2, src.to(torch.float32).reshape(B, 1, Tt).contiguous(), self.w.pitch_emb_w,
                             self.w.pitch_emb_b, in_lens32)
        energy_pred = None
        if energy_tgt is None:
            energy_pred = self._pred_fwd(enc2, in_lens32, self.pred["energy"], None)
            esrc = energy_pred.view(B, 1, Tt)
        else:
            esrc = energy_tgt
        enc3 = enc2.clone()
        ops.scalar_conv_add_(enc3, esrc.to(torch.float32).reshape(B, 1, Tt).contiguous(), self.w.energy_emb_w,
                             self.w.energy_emb_b, in_lens32)
        durs = (dur_pred if dur_tgt is None else dur_tgt).to(torch.float32).contiguous()
        cum, dec_lens = ops.duration_scan(durs, pace, None)       # regulate_len(..., mel_max_len=None), model.py:473-475
        T_out = int(dec_lens.max().item())
        if T_out <= 0:
            raise ValueError("infer: every predicted duration rounds to zero frames")
        regulated = ops.regulate_gather(enc3, cum, T_out)
        y = ops.embed_pos(None, None, regulated, dec_lens, self.inv_freq, B, T_out, D_MODEL)
        for L in self.dec_layers:
            y = self._layer_fwd(y, dec_lens, L, None)
        mel_out = ops.conv_fwd(y, self.w.proj_w, bias=self.w.proj_b)
        return mel_out.permute(0, 2, 1), dec_lens.to(torch.int64), dur_pred, pitch_pred, energy_pred

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, criterion, scale=1.0, grad_sync=None, kl=None):
        """Reverse pass for the loss ``criterion`` just evaluated on this module's last forward() output. Gradients of
        the stage's trainable parameters are ACCUMULATED into the gradient arena (zero_grad() clears it), scaled by
        ``scale`` (the 1/gam of xva_train.py:806). ``grad_sync`` (parallel.GradSync) is told which arena slices are
        final as the pass proceeds, so their all-reduce overlaps the rest of the backward. Stage 1 only: ``kl`` =
        (AttentionBinarizationLoss already evaluated on this forward's attn_hard / attn_soft, kl_weight) adds the
        binarization term of xva_train.py:792-798 to the loss being differentiated."""
        ctx = self._ctx
        if grad_sync is not None:
            scale = scale * grad_sync.loss_scale
        def ready(*prefixes, flush=False):
            if grad_sync is not None:
                self._join_side()          # the slice must be final on the stream the all-reduce is ordered after
                grad_sync.ready(list(prefixes), flush)

        if ctx is None:
            raise RuntimeError("backward() needs a forward() in training mode first")
        stage = ctx.stage
        if stage == 1:
            return self._backward_stage1(criterion, scale, kl, ready)
        lens = ctx.in_lens
        B, Tt = ctx.B, ctx.Tt
        seeds = criterion.grad_seeds(scale)
        if stage == 2:
            d_enc = self._pred_bwd(seeds["log_dur"], lens, self.pred["duration"], ctx.preds["duration"])
            ready("duration_predictor", flush=True)
        else:
            # projection
            dmel = seeds["mel"]                                   # [B,T_out,96], columns 80.. are zero
            dm = dmel[..., :N_MEL]
            T_out = ctx.T_out
            # the predictors' backward needs only its loss seeds: with the streams on it is ordered after THIS point
            # only (an event), so inside a captured graph it runs next to the decoder's backward although it is issued
            # after it -- program order, launches and arguments do not depend on the switch
            pred_par = self._side_on() and stage == 3 and os.environ.get("XVA_PRED_STREAM", "1") != "0"
            if pred_par:
                ps = self._pred_side()
                ev_start = torch.cuda.Event()
                ev_start.record(torch.cuda.current_stream())
            ops.conv_wgrad(dm, ctx.dec_out, (0,), out=self.g.proj_w, accumulate=True)
            ops.colsum_(B * T_out, N_MEL, dmel.shape[2], dmel, self.g.proj_b)
            dy = ops.conv_dgrad(dm, self.w.proj_w, lens=ctx.dec_lens)
            with ops.nvtx("fastpitch.bwd.decoder"):
                for i, (L, c) in enumerate(zip(reversed(self.dec_layers), reversed(ctx.dec))):
                    dy = self._layer_bwd(dy, ctx.dec_lens, L, c)
                    ready(f"decoder.layers.{N_LAYERS - 1 - i}")
            d_enc = ops.regulate_scatter(dy, ctx.cum, Tt)         # pos-emb has no parameters: dy is d(regulated)
            del dy
            ops.scalar_conv_bwd_(d_enc, ctx.energy_tgt, self.g.energy_emb_w, self.g.energy_emb_b)
            if stage == 3:
                if pred_par:
                    ps.wait_event(ev_start)
                    with torch.cuda.stream(ps):
                        d_en = self._pred_bwd(seeds["energy"], lens, self.pred["energy"], ctx.preds["energy"])
                        d_pi = self._pred_bwd(seeds["pitch"], lens, self.pred["pitch"], ctx.preds["pitch"])
                    cur = torch.cuda.current_stream()
                    cur.wait_stream(ps)
                    d_en.record_stream(cur)
                    d_pi.record_stream(cur)
                else:
                    d_en = self._pred_bwd(seeds["energy"], lens, self.pred["energy"], ctx.preds["energy"])
                    d_pi = self._pred_bwd(seeds["pitch"], lens, self.pred["pitch"], ctx.preds["pitch"])
                d_enc.add_(d_en)          # d(enc2) = d(enc3) + the energy predictor's input gradient
                ops.scalar_conv_bwd_(d_enc, ctx.pitch_tgt, self.g.pitch_emb_w, self.g.pitch_emb_b)
                d_enc.add_(d_pi)
            # the tail of the arena (pitch/energy predictors + embeddings, proj) is final, touched or not
            ready("pitch_predictor", "pitch_emb", "energy_predictor", "energy_emb", "proj", flush=True)
        with ops.nvtx("fastpitch.bwd.encoder"):
            for i, (L, c) in enumerate(zip(reversed(self.enc_layers), reversed(ctx.enc))):
                d_enc = self._layer_bwd(d_enc, lens, L, c)
                if i < N_LAYERS - 1:
                    ready(f"encoder.layers.{N_LAYERS - 1 - i}")
        ops.embed_bwd_(ctx.tokens, d_enc, self.g.emb)
        ready("encoder.layers.0", "encoder.word_emb", flush=True)
        self._join_side()
        self._ctx = None

    def step_dropout(self):
        """Advance the device-side dropout counter (one increment per optimizer micro-step)."""
        ops.counter_add_(self.step_counter, 1)


class FastPitchLoss:
    """FastPitchLoss, fastpitch/loss_function.py:29-154, for training stages 2-4: masked MSE terms reduced on the device
    in fp64. ``forward`` returns (loss, meta) as 0-dim device tensors (no host sync); ``grad_seeds`` hands the
    d(loss)/d(prediction) tensors to FastPitch.backward."""

    def __init__(self, dur_predictor_loss_scale=0.1, pitch_predictor_loss_scale=0.1, attn_loss_scale=1.0,
                 energy_predictor_loss_scale=0.1, gpus=None):
        self.dur_predictor_loss_scale = dur_predictor_loss_scale
        self.pitch_predictor_loss_scale = pitch_predictor_loss_scale
        self.energy_predictor_loss_scale = energy_predictor_loss_scale
        self.attn_loss_scale = attn_loss_scale
        self.training_stage = 3
        self._saved = None
        self.world, self.group = 1, None

    def set_distributed(self, world, group=None):
        """One process per GPU: every masked MSE becomes sum_global(err * mask) / sum_global(mask) -- what the
        reference computes on the outputs nn.DataParallel gathered (xva_train.py:790, loss_function.py:90-117), and what
        one GPU running the global batch computes -- instead of the mean of per-rank ratios: the {sum, count} pairs are
        all-reduced (8 doubles) before the ratios and before grad_seeds() divides by the counts. Each rank then holds
        d(global loss)/d(its own predictions), so the gradient all-reduce is a plain SUM (GradSync(mean=False)). The
        stage-1 CTC term is a mean over utterances (equal batch per rank): mean of means, scaled by 1/world."""
        self.world, self.group = int(world), group
        return self

    def _all_reduce(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def __call__(self, model_out, targets, is_training=True, meta_agg="mean", training_stage=None):
        return self.forward(model_out, targets, is_training, meta_agg, training_stage)

    def forward(self, model_out, targets, is_training=True, meta_agg="mean", training_stage=None):
        """``training_stage`` (the reference passes it per call, loss_function.py:63) overrides the attribute."""
        (mel_out, dec_mask, dur_pred, log_dur_pred, pitch_pred, pitch_tgt, energy_pred, energy_tgt, _, _, attn_dur,
         attn_logprob, input_lens) = model_out
        mel_tgt, in_lens, out_lens, max_inp_lengths = targets[:4]
        if training_stage is not None:
            self.training_stage = int(training_stage)
        stage = self.training_stage
        dev = input_lens.device
        lens32 = input_lens.to(torch.int32)
        if stage == 1:
            # AttentionCTCLoss over the whole batch in one launch (loss_function.py:74-82); the kernel also leaves
            # d(mean cost)/d(attn_logprob), which grad_seeds() hands to FastPitch.backward
            lp = attn_logprob.reshape(attn_logprob.shape[0], attn_logprob.shape[-2], attn_logprob.shape[-1])
            cost, gctc = ops.attn_ctc(lp, lens32.contiguous(), out_lens.to(torch.int32).contiguous())
            attn_loss = cost.mean()
            if self.world > 1:
                attn_loss = self._all_reduce(attn_loss.clone()) / self.world
            loss = attn_loss * self.attn_loss_scale
            self._saved = {"stage": 1, "gctc": gctc}
            return loss, {"loss": loss, "attn_loss": attn_loss}
        acc = torch.zeros(4, 2, device=dev, dtype=torch.float64)   # rows: mel, dur, pitch, energy = {sum sq err, count}
        zero = torch.zeros((), device=dev, dtype=torch.float64)
        mel_loss = dur_loss = pitch_loss = energy_loss = zero
        saved = {"stage": stage, "acc": acc, "lens": lens32}
        if stage == 2:
            tgt = attn_dur.to(torch.float32).contiguous()
            ops.lens_mse(log_dur_pred, tgt, lens32, acc[1], log1p_tgt=True)
            self._all_reduce(acc)
            dur_loss = acc[1, 0] / acc[1, 1]
            saved.update(log_dur_pred=log_dur_pred, dur_tgt=tgt)
        else:
            mt = mel_tgt.to(torch.float32).contiguous()
            ops.mel_mse(mel_out, mt, acc[0])
            saved.update(mel_out=mel_out, mel_tgt=mt)
            if stage == 3:
                pp, pt = pitch_pred.reshape(pitch_pred.shape[0], -1), pitch_tgt.reshape(pitch_tgt.shape[0], -1)
                ops.lens_mse(pp, pt, lens32, acc[2])
                ops.lens_mse(energy_pred, energy_tgt, lens32, acc[3])
                saved.update(pitch_pred=pp, pitch_tgt=pt, energy_pred=energy_pred, energy_tgt=energy_tgt)
            self._all_reduce(acc)                      # global {sum, count} of every term, one 8-double message
            mel_loss = acc[0, 0] / acc[0, 1]
            if stage == 3:
                pitch_loss = acc[2, 0] / acc[2, 1]
                energy_loss = acc[3, 0] / acc[3, 1]
        loss = (mel_loss + dur_loss * self.dur_predictor_loss_scale + pitch_loss * self.pitch_predictor_loss_scale
                + energy_loss * self.energy_predictor_loss_scale)
        self._saved = saved
        meta = {"loss": loss, "mel_loss": mel_loss, "duration_predictor_loss": dur_loss, "pitch_loss": pitch_loss,
                "energy_loss": energy_loss}
        return loss, meta

    def grad_seeds(self, scale=1.0):
        s = self._saved
        if s is None:
            raise RuntimeError("FastPitchLoss.grad_seeds() needs a forward() first")
        if s["stage"] == 1:
            return {"attn_logprob": (s["gctc"], scale * self.attn_loss_scale / self.world)}
        acc, lens = s["acc"], s["lens"]
        out = {}
        if s["stage"] == 2:
            out["log_dur"] = ops.lens_mse_grad(s["log_dur_pred"], s["dur_tgt"], lens, acc[1],
                                               scale * self.dur_predictor_loss_scale, log1p_tgt=True)
        else:
            out["mel"] = ops.mel_mse_grad(s["mel_out"], s["mel_tgt"], acc[0], scale, 96)
            if s["stage"] == 3:
                out["pitch"] = ops.lens_mse_grad(s["pitch_pred"], s["pitch_tgt"], lens, acc[2],
                                                 scale * self.pitch_predictor_loss_scale)
                out["energy"] = ops.lens_mse_grad(s["energy_pred"], s["energy_tgt"], lens, acc[3],
                                                  scale * self.energy_predictor_loss_scale)
        return out


class AttentionBinarizationLoss:
    """AttentionBinarizationLoss, fastpitch/attn_loss_function.py:47-54: -sum_{hard == 1} log(clamp(soft, eps)) / sum(hard),
    reduced on the device in fp64. The trainer adds kl_weight times it to the stage-1 loss (xva_train.py:792-798); pass
    ``kl=(this, kl_weight)`` to FastPitch.backward for its gradient."""

    def __init__(self):
        self._saved = None
        self.world, self.group = 1, None

    def set_distributed(self, world, group=None):
        """Global {sum, count} across ranks, as FastPitchLoss.set_distributed."""
        self.world, self.group = int(world), group
        return self

    def __call__(self, hard_attention, soft_attention, eps=1e-12):
        return self.forward(hard_attention, soft_attention, eps)

    def forward(self, hard_attention, soft_attention, eps=1e-12):
        hard = hard_attention.to(torch.float32).contiguous()
        soft = soft_attention.to(torch.float32).contiguous()
        acc = torch.zeros(2, device=soft.device, dtype=torch.float64)
        ops.attn_bin_loss(hard, soft, acc, eps)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        self._saved = (hard, soft, acc, float(eps))
        return -acc[0] / acc[1]

    def saved(self):
        """(hard, soft, acc) of the last forward, for ops.attn_grad_combine."""
        if self._saved is None:
            raise RuntimeError("AttentionBinarizationLoss.saved() needs a forward() first")
        return self._saved[:3]


class Lamb:
    """Lamb, lamb.py:8-106, as one multi-tensor launch pair over the model's flat arena, preceded by the
    clip_grad_norm_(max_norm) the trainer applies (xva_train.py:857). ``param_groups[0]['lr']`` is what
    adjust_learning_rate (xva_train.py:1252-1261) writes."""

    CHUNK = 8192

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, clip_grad_norm=1000.0):
        self.model = model
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.clip = clip_grad_norm
        A = model.arena
        if A.m is None:
            A.m = torch.zeros_like(A.p)
            A.v = torch.zeros_like(A.p)
        self.lr_dev = torch.zeros(1, device=A.p.device, dtype=torch.float32)
        self.lr_on_device = False   # True: lr_dev is maintained by the caller (CUDA-graph replay), step() leaves it
        self._tables = {}
        self.steps = 0

    def _table(self, stage):
        if stage not in self._tables:
            A = self.model.arena
            keys = trainable_keys(stage)
            rec = []
            for ti, k in enumerate(keys):
                off, n = A.offset[k], int(math.prod(A.pshape[k]))
                for s in range(0, n, self.CHUNK):
                    rec.append((off + s, min(self.CHUNK, n - s), ti))
            import numpy as np
            arr = np.zeros(len(rec), dtype=np.dtype([("start", "<i8"), ("len", "<i4"), ("tensor", "<i4")]))
            for i, (a, b, c) in enumerate(rec):
                arr[i] = (a, b, c)
            chunks = torch.from_numpy(arr.view(np.uint8).copy()).to(A.p.device)
            self._tables[stage] = (chunks, len(rec), len(keys))
        return self._tables[stage]

    def zero_grad(self, set_to_none=True):
        self.model.zero_grad()

    def step(self, closure=None):
        with ops.nvtx("fastpitch.lamb"):
            return self._step()

    def _step(self):
        A = self.model.arena
        g = self.param_groups[0]
        chunks, n_chunks, n_tensors = self._table(self.model.training_stage)
        if not self.lr_on_device:
            self.lr_dev.fill_(float(g["lr"]))
        scratch = torch.zeros(2 * n_tensors + 1, device=A.p.device, dtype=torch.float64)
        gn = scratch[2 * n_tensors:]
        if self.clip is not None and self.clip > 0:
            ops.grad_sqnorm(A.g, chunks, n_chunks, gn)
        ops.lamb_step(A.p, A.g, A.m, A.v, chunks, n_chunks, scratch, gn if self.clip else None,
                      self.clip if self.clip else 0.0, self.lr_dev, g["betas"][0], g["betas"][1], g["eps"],
                      g["weight_decay"], p_tf32=A.w)
        self.steps += 1
        self.last_grad_sqnorm = gn

    def state_dict(self):
        """{'state': {reference key: {'exp_avg', 'exp_avg_sq'}}, 'param_groups'} with tensors in the reference's shapes."""
        A = self.model.arena
        st = OrderedDict()
        for k, _ in A.spec:
            st[k] = {"step": self.steps, "exp_avg": _to_ref(k, A.view(A.m, k), A.rshape[k]),
                     "exp_avg_sq": _to_ref(k, A.view(A.v, k), A.rshape[k])}
        return {"state": st, "param_groups": [dict(g) for g in self.param_groups], "steps": self.steps}

    def load_state_dict(self, sd):
        A = self.model.arena
        for k, st in sd.get("state", {}).items():
            if k in A.offset:
                A.view(A.m, k).copy_(_to_packed(k, st["exp_avg"].to(torch.float32)))
                A.view(A.v, k).copy_(_to_packed(k, st["exp_avg_sq"].to(torch.float32)))
        self.steps = int(sd.get("steps", 0))
        if sd.get("param_groups"):
            self.param_groups[0].update(sd["param_groups"][0])


def adjust_learning_rate(total_iter, opt, learning_rate, warmup_iters=None):
    """xva_train.py:1252-1261 (noam)."""
    if warmup_iters == 0:
        scale = 1.0
    elif total_iter > warmup_iters:
        scale = 1.0 / (total_iter ** 0.5)
    else:
        scale = total_iter / (warmup_iters ** 1.5)
    for param_group in opt.param_groups:
        param_group["lr"] = learning_rate * scale
