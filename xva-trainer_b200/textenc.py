"""xVAPitch text encoder and pitch predictor on the B200 engine.

    RelativePositioningPitchEnergyEncoder
                  drop-in for python/xvapitch/model.py:1268 as xVAPitch builds its pitch predictor (model.py:154-168): the
                  same transformer layers over cat(text-encoder output, speaker embedding) = 708 / 780 channels, 3 layers,
                  a 1-channel projection (class docstring below)
    TextEncoder   drop-in for python/xvapitch/model.py:1089 ``TextEncoder`` (same constructor arguments, ``forward`` with
                  both of its calls -- stats=False: embedding + language embedding + transformer; stats=True: the 1x1
                  projection to the prior's mean / log-scale --, same state_dict keys and shapes) over
                  python/xvapitch/glow_tts.py:373-485 ``RelativePositionTransformer`` (relative-position multi-head
                  attention with a window of 4, glow_tts.py:59-292; kernel-3 conv FFN, :324-372; LayerNorm2, :34-56).

Every product (the q / k / v / o projections, q.k^T, q.E_k^T, P.v, P_band.E_v, the two FFN convolutions and all their
input and weight gradients) is a launch of the tcgen05 tap-GEMM of libxva_b200.so on channels-last [B, T, C] tensors; the
softmax, LayerNorm and column-sum kernels are the ones the FastPitch FFT block uses; the steps that are new here --
embedding + language concat, adding the relative-position logits onto the score band, reading the band of the attention
weights back in relative indexing, padded operand copies -- are csrc/relattn.cu.

Layout decisions (all forced by operand alignment, none changes a result):
  * C = hidden + language channels is 196 / 204 / 268 in the reference's configurations, so a head is 98 / 102 / 134 columns:
    not a multiple of 4 floats, i.e. head h of a packed q | k | v row would not start on the 16 bytes a TMA operand base
    needs. Each head is padded to dkp = 128 (160) columns in the projection's OUTPUT (zero weight rows and biases), in
    E_k / E_v and in the input columns of conv_o; pad entries have zero gradients (every product that reaches them has a
    zero factor), so AdamW leaves them zero.
  * weights whose input dimension is C are stored with a row pitch of Cp = C rounded up to 32: the input-gradient GEMM
    reads them MN-major in 32-column chunks. Activations are [B, T, C] contiguous (what the LayerNorm kernels take); the
    four tensors per layer that a weight-gradient GEMM reads MN-major get a padded copy (xva_pad_cols) in the backward.
  * all parameters live in ONE flat fp32 tensor in the layout the kernels read (``flat``; its ``.grad`` is the gradient
    arena the weight-gradient GEMMs accumulate into), with a tf32-rounded copy refreshed by one launch per forward.
    state_dict() / load_state_dict() convert to and from the reference's keys and shapes.
There is no CPU path: the module raises without the library.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import capi, ops

REL_WINDOW = 4           # model.py:1137 rel_attn_window_size=4
REL_COLS = 32            # the 2 W + 1 = 9 relative positions padded to one 32-column operand chunk


def _need_cuda(device):
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if dev.type != "cuda":
        raise capi.XvaError("the xVAPitch modules (B200 build) need a CUDA device: there is no CPU path")
    capi.load()
    capi.call("xva_device_check", dev.index or 0)
    return dev


def _up(n, m):
    return (n + m - 1) // m * m


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _RelTransformer(nn.Module):
    """What TextEncoder and RelativePositioningPitchEnergyEncoder share: the layers of RelativePositionTransformer
    (python/xvapitch/glow_tts.py:373-485) on the engine -- dimensions and padded layouts, the flat parameter arena, and the
    forward / backward of the attention half (glow_tts.py:159-214 + LayerNorm, :476-478) and of the FFN half (:353-357 +
    LayerNorm, :480-482) of one layer."""

    def _setup(self, C, num_heads, hidden_channels_ffn, kernel_size, dropout_p, max_c=1024):
        if C % num_heads or C % 4 or C > max_c:
            raise NotImplementedError(f"{C} channels: need a multiple of 4 and of the head count, at most {max_c} (LayerNorm kernels)")
        if hidden_channels_ffn % 32:
            raise NotImplementedError("FFN channels must be a multiple of 32")
        self.num_heads, self.hidden_channels_ffn, self.kernel_size = int(num_heads), int(hidden_channels_ffn), int(kernel_size)
        self.dropout_p = float(dropout_p)
        self.C, self.Cp = C, _up(C, 32)
        self.dk = C // self.num_heads
        self.dkp = _up(self.dk, 32)
        self.W = REL_WINDOW
        self.shifts = tuple(j - (self.kernel_size - 1) // 2 for j in range(self.kernel_size))    # glow_tts.py:354-361

    def _layer_spec(self, i, ffn=True):
        """(name, kernel-layout shape) of layer i's tensors in the arena."""
        C, H, dkp, Cp, F, k = self.C, self.num_heads, self.dkp, self.Cp, self.hidden_channels_ffn, self.kernel_size
        spec = [(f"l{i}.qkv_w", (3 * H * dkp, Cp)), (f"l{i}.qkv_b", (3 * H * dkp,)), (f"l{i}.ek", (REL_COLS, dkp)),
                (f"l{i}.ev", (REL_COLS, dkp)), (f"l{i}.o_w", (C, H * dkp)), (f"l{i}.o_b", (C,)), (f"l{i}.ln1_g", (C,)),
                (f"l{i}.ln1_b", (C,))]
        if ffn:
            spec += [(f"l{i}.w1", (k, F, Cp)), (f"l{i}.b1", (F,)), (f"l{i}.w2", (k, C, F)), (f"l{i}.b2", (C,)),
                     (f"l{i}.ln2_g", (C,)), (f"l{i}.ln2_b", (C,))]
        return spec

    def _alloc(self, spec, dev, seed):
        """One flat fp32 parameter tensor (every entry on a 64-float boundary), its gradient arena and tf32 copy."""
        self._spec, self._off, n = spec, {}, 0
        for name, shape in spec:
            self._off[name] = n
            n += _up(int(math.prod(shape)), 64)
        self.flat = nn.Parameter(torch.zeros(n, device=dev, dtype=torch.float32))
        self.flat.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self._w = torch.zeros(n, device=dev, dtype=torch.float32)          # tf32-rounded copy: what the GEMMs read
        self.seed = int(seed)
        self.step_counter = torch.zeros(1, device=dev, dtype=torch.int64)   # device-side dropout counter
        self._site = 0
        self._ctx = None

    def _views(self, flat):
        """{name: view of `flat` in the kernel layout} (flat: the parameters, their gradients or the rounded copy)."""
        out = {}
        for name, shape in self._spec:
            o = self._off[name]
            out[name] = flat[o:o + int(math.prod(shape))].view(shape)
        return out

    def _layer(self, V, i):
        """Per-layer operand views of an arena view dict: weights as [taps, N, K] with K cut to the real channel count."""
        C = self.C
        g = lambda n: V[f"l{i}.{n}"]
        ns = _NS(qkv_w=g("qkv_w")[None, :, :C], qkv_b=g("qkv_b"), ek=g("ek")[None], ev=g("ev")[None], o_w=g("o_w")[None],
                 o_b=g("o_b"), ln1_g=g("ln1_g"), ln1_b=g("ln1_b"))
        if f"l{i}.w1" in V:
            ns.__dict__.update(w1=g("w1")[..., :C], b1=g("b1"), w2=g("w2"), b2=g("b2"), ln2_g=g("ln2_g"), ln2_b=g("ln2_b"))
        return ns

    @property
    def arena(self):
        """What parallel.GradSync reads: the gradient arena ``g``, the parameters ``p`` and where each tensor
        lies. Every gradient of the module is in the one flat buffer, so the data-parallel exchange is all-reduces of its
        slices -- ``GradSync(module, world).ready(["l2"])`` as soon as backward has finished layer 2, ``finish()`` before
        the optimizer reads it; the pad entries stay zero on every rank (0 + 0)."""
        return _NS(g=self.flat.grad, p=self.flat.data, offset=self._off, pshape=dict(self._spec))

    def to(self, *args, **kwargs):
        """The device is fixed at construction (arena, gradient arena, operand copy and dropout counter live there)."""
        return self

    def zero_grad(self, set_to_none=False):
        if self.flat.grad is None:
            self.flat.grad = torch.zeros_like(self.flat.data)
        else:
            self.flat.grad.zero_()

    def _drop(self):
        """(p, seed) of the next dropout site of this pass, numbered in call order (the backward re-derives the masks)."""
        self._site += 1
        p = self.dropout_p if self.training else 0.0
        return p, (self.seed * 0x9E3779B1 + self._site * 0x85EBCA77) & 0xFFFFFFFFFFFF

    def step_dropout(self):
        """Advance the device-side dropout counter: call once per optimizer micro-step (fresh masks on a graph replay)."""
        ops.counter_add_(self.step_counter, 1)

    def _weights(self):
        ops.round_tf32_(self.flat.data, self._w)
        return self._views(self._w), self._views(self.flat.data)

    def _to_ref(self, V, params=True):
        """Arena views (kernel layout, padded) -> {reference key: tensor in the reference's shape} (copies). params=False:
        V holds gradients (parameters that are not on the path have none)."""
        C, H, dk, dkp, NR = self.C, self.num_heads, self.dk, self.dkp, 2 * self.W + 1
        out = OrderedDict()
        for i in range(self.num_layers):
            a, f = f"encoder.attn_layers.{i}", f"encoder.ffn_layers.{i}"
            qw = V[f"l{i}.qkv_w"].view(3, H, dkp, self.Cp)[:, :, :dk, :C]
            qb = V[f"l{i}.qkv_b"].view(3, H, dkp)[:, :, :dk]
            out[f"{a}.emb_rel_k"] = V[f"l{i}.ek"][:NR, :dk].clone()[None]
            out[f"{a}.emb_rel_v"] = V[f"l{i}.ev"][:NR, :dk].clone()[None]
            for s, n in enumerate("qkv"):
                out[f"{a}.conv_{n}.weight"] = qw[s].reshape(C, C, 1).clone()
                out[f"{a}.conv_{n}.bias"] = qb[s].reshape(C).clone()
            out[f"{a}.conv_o.weight"] = V[f"l{i}.o_w"].view(C, H, dkp)[:, :, :dk].reshape(C, C, 1).clone()
            out[f"{a}.conv_o.bias"] = V[f"l{i}.o_b"].clone()
            out[f"encoder.norm_layers_1.{i}.gamma"] = V[f"l{i}.ln1_g"].clone()
            out[f"encoder.norm_layers_1.{i}.beta"] = V[f"l{i}.ln1_b"].clone()
            if f"l{i}.w1" not in V:
                continue                                   # a layer whose FFN half is not on the path (_extra_to_ref fills it in)
            out[f"{f}.conv_1.weight"] = V[f"l{i}.w1"][..., :C].permute(1, 2, 0).contiguous()
            out[f"{f}.conv_1.bias"] = V[f"l{i}.b1"].clone()
            out[f"{f}.conv_2.weight"] = V[f"l{i}.w2"].permute(1, 2, 0).contiguous()
            out[f"{f}.conv_2.bias"] = V[f"l{i}.b2"].clone()
            out[f"encoder.norm_layers_2.{i}.gamma"] = V[f"l{i}.ln2_g"].clone()
            out[f"encoder.norm_layers_2.{i}.beta"] = V[f"l{i}.ln2_b"].clone()
        self._extra_to_ref(V, out, params)
        return OrderedDict((k, out[k]) for k, _ in self._ref_spec() if k in out)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = OrderedDict() if destination is None else destination
        for k, v in self._to_ref(self._views(self.flat.detach())).items():
            out[prefix + k] = v
        return out

    def grads(self):
        """{reference key: gradient in the reference's shape} for every parameter on the path."""
        return self._to_ref(self._views(self.flat.grad), params=False)

    def load_state_dict(self, state_dict, strict=True):
        known = dict(self._ref_spec())
        missing = [k for k in known if k not in state_dict]
        unexpected = [k for k in state_dict if k not in known]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        C, H, dk, dkp, NR = self.C, self.num_heads, self.dk, self.dkp, 2 * self.W + 1
        V = self._views(self.flat.data)
        get = lambda k: None if k not in state_dict else state_dict[k].detach().to(device=self.flat.device, dtype=torch.float32)
        with torch.no_grad():
            for k, t in state_dict.items():
                if k in known and tuple(t.shape) != tuple(known[k]):
                    raise RuntimeError(f"load_state_dict: {k} has shape {tuple(t.shape)}, expected {tuple(known[k])}")

            def put(dst, key, fn=lambda t: t):
                t = get(key)
                if t is not None:
                    dst.copy_(fn(t))

            for i in range(self.num_layers):
                a, f = f"encoder.attn_layers.{i}", f"encoder.ffn_layers.{i}"
                qw = V[f"l{i}.qkv_w"].view(3, H, dkp, self.Cp)
                qb = V[f"l{i}.qkv_b"].view(3, H, dkp)
                put(V[f"l{i}.ek"][:NR, :dk], f"{a}.emb_rel_k", lambda t: t[0])
                put(V[f"l{i}.ev"][:NR, :dk], f"{a}.emb_rel_v", lambda t: t[0])
                for s, n in enumerate("qkv"):
                    put(qw[s, :, :dk, :C], f"{a}.conv_{n}.weight", lambda t: t[:, :, 0].view(H, dk, C))
                    put(qb[s, :, :dk], f"{a}.conv_{n}.bias", lambda t: t.view(H, dk))
                put(V[f"l{i}.o_w"].view(C, H, dkp)[:, :, :dk], f"{a}.conv_o.weight", lambda t: t[:, :, 0].view(C, H, dk))
                put(V[f"l{i}.o_b"], f"{a}.conv_o.bias")
                put(V[f"l{i}.ln1_g"], f"encoder.norm_layers_1.{i}.gamma")
                put(V[f"l{i}.ln1_b"], f"encoder.norm_layers_1.{i}.beta")
                if f"l{i}.w1" not in V:
                    continue
                put(V[f"l{i}.w1"][..., :C], f"{f}.conv_1.weight", lambda t: t.permute(2, 0, 1))
                put(V[f"l{i}.b1"], f"{f}.conv_1.bias")
                put(V[f"l{i}.w2"], f"{f}.conv_2.weight", lambda t: t.permute(2, 0, 1))
                put(V[f"l{i}.b2"], f"{f}.conv_2.bias")
                put(V[f"l{i}.ln2_g"], f"encoder.norm_layers_2.{i}.gamma")
                put(V[f"l{i}.ln2_b"], f"encoder.norm_layers_2.{i}.beta")
            self._extra_from_ref(V, put, get)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def reset_parameters(self, seed=1234):
        """The reference constructors' initialisation, seeded and generated on the CPU (identical on every rank): emb ~
        N(0, hidden^-1/2) (model.py:1119), emb_rel ~ N(0, d_k^-1/2), xavier-uniform conv_q / conv_k / conv_v
        (glow_tts.py:139-157), torch's Conv1d default (U(+-1/sqrt(fan_in))) elsewhere, LayerNorm2 gamma = 1 / beta = 0."""
        gen = torch.Generator().manual_seed(int(seed))
        sd = OrderedDict()
        shapes = dict(self._ref_spec())
        for key, shape in self._ref_spec():
            if key == "emb.weight":
                t = torch.randn(shape, generator=gen) * shape[1] ** -0.5
            elif "emb_rel" in key:
                t = torch.randn(shape, generator=gen) * self.dk ** -0.5
            elif key.endswith("gamma"):
                t = torch.ones(shape)
            elif key.endswith("beta"):
                t = torch.zeros(shape)
            elif key.endswith(("conv_q.weight", "conv_k.weight", "conv_v.weight")):
                bound = math.sqrt(6.0 / (shape[0] + shape[1]))
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            elif key.endswith(".weight"):
                bound = 1.0 / math.sqrt(shape[1] * shape[2])
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            else:                                         # a convolution's bias: the same fan-in bound as its weight
                wshape = shapes[key[:-4] + "weight"]
                bound = 1.0 / math.sqrt(wshape[1] * wshape[2])
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            sd[key] = t
        self.load_state_dict(sd)

    # ------------------------------------------------------------------------------------------ one layer, forward
    def _attn_fwd(self, x, lens, lens_rep, L, Lf):
        """x [B, T, C] (zero rows past the lengths) -> (LayerNorm1(x + dropout(attention(x))), what the backward needs)."""
        B, T, _ = x.shape
        H, dkp, Wd, sd = self.num_heads, self.dkp, self.W, self.step_counter
        Tp = _up(T, 32)
        alpha = 1.0 / math.sqrt(self.dk)
        qkv = ops.conv_fwd(x, L.qkv_w, bias=Lf.qkv_b, round_out=True)                     # [B, T, 3 H dkp]
        head = lambda s, h: qkv[..., (s * H + h) * dkp:(s * H + h + 1) * dkp]
        s_ = torch.empty(H, B, T, Tp, device=x.device, dtype=torch.float32)
        rel = torch.empty(H, B, T, REL_COLS, device=x.device, dtype=torch.float32)
        for h in range(H):
            ops.bmm_nt(head(0, h), head(1, h), alpha=alpha, out=s_[h][..., :T])          # q.k^T / sqrt(d_k)
            ops.conv_fwd(head(0, h), L.ek, out=rel[h], alpha=alpha)                       # q.E_k^T / sqrt(d_k)
        ops.rel_band_add_(s_.view(H * B, T, Tp), rel.view(H * B, T, REL_COLS), T, Wd)
        p_att, seed_att = self._drop()
        P, Pd = ops.softmax_fwd(s_.view(H * B, T, Tp), lens_rep, T, p_att, seed_att, sd)
        del s_, rel
        PB = ops.rel_band_gather(Pd, T, Wd, REL_COLS)                                     # [H B, T, 32]
        vec = torch.empty(B, T, H * dkp, device=x.device, dtype=torch.float32)
        Pd4, PB4 = Pd.view(H, B, T, Tp), PB.view(H, B, T, REL_COLS)
        for h in range(H):
            pv = ops.bmm_nn(Pd4[h][..., :T], head(2, h))                                  # P.v
            ops.conv_dgrad(PB4[h], L.ev, out=vec[..., h * dkp:(h + 1) * dkp], residual=pv, round_out=True)  # + P_band.E_v
        p1, seed1 = self._drop()
        pre1 = ops.conv_fwd(vec, L.o_w, bias=Lf.o_b, residual=x, drop_p=p1, seed=seed1, seed_dev=sd)
        y1, sv1 = ops.layernorm_fwd(pre1, Lf.ln1_g, Lf.ln1_b, lens)
        return y1, _NS(x=x, qkv=qkv, P=P, Pd=Pd, PB=PB, vec=vec, sv1=sv1, y1=y1, att=(p_att, seed_att), d1=(p1, seed1))

    def _ffn_fwd(self, y1, lens, L, Lf):
        """y1 -> (LayerNorm2(y1 + dropout(conv_2(dropout(relu(conv_1(y1))) * mask) * mask)), saved)."""
        sd = self.step_counter
        pf, seedf = self._drop()
        hdn = ops.conv_fwd(y1, L.w1, self.shifts, bias=Lf.b1, relu=True, lens=lens, drop_p=pf, seed=seedf, seed_dev=sd,
                           round_out=True)
        p2, seed2 = self._drop()
        pre2 = ops.conv_fwd(hdn, L.w2, self.shifts, bias=Lf.b2, residual=y1, drop_p=p2, seed=seed2, seed_dev=sd)
        y2, sv2 = ops.layernorm_fwd(pre2, Lf.ln2_g, Lf.ln2_b, lens)
        return y2, _NS(h=hdn, sv2=sv2, df=(pf, seedf), d2=(p2, seed2))

    # ------------------------------------------------------------------------------------------ one layer, backward
    def _ffn_bwd(self, dy, s, lens, L, Lf, g):
        """dy = dL/d(layer output) -> dL/dy1 (FeedForwardNetwork + LayerNorm2, glow_tts.py:353-357, 481-482)."""
        B, T, C = dy.shape
        Cp, F, sd = self.Cp, self.hidden_channels_ffn, self.step_counter
        dx2, dbr2 = ops.layernorm_bwd(dy, s.sv2, Lf.ln2_g, lens, g.ln2_g, g.ln2_b, dbias=g.b2, want_drop=True,
                                      drop_pre_p=s.d2[0], seed_pre=s.d2[1], seed_dev=sd)
        ops.conv_wgrad(ops.pad_cols(dbr2, Cp)[..., :C], s.h, self.shifts, out=g.w2, accumulate=True)
        dh = ops.conv_dgrad(dbr2, L.w2, self.shifts, gate=s.h, drop_p=s.df[0], seed=s.df[1], seed_dev=sd, round_out=True)
        ops.conv_wgrad(dh, ops.pad_cols(s.y1, Cp)[..., :C], self.shifts, out=g.w1, accumulate=True)
        ops.colsum_(B * T, F, F, dh, g.b1)
        return ops.conv_dgrad(dh, L.w1, self.shifts, residual=dx2)

    def _attn_bwd(self, dy1, s, lens, L, Lf, g):
        """dy1 = dL/dy1 -> dL/d(layer input) (RelativePositionMultiHeadAttention + LayerNorm1, glow_tts.py:159-214, 476-478)."""
        B, T, C = dy1.shape
        Cp, H, dkp, Wd, sd = self.Cp, self.num_heads, self.dkp, self.W, self.step_counter
        Tp = _up(T, 32)
        alpha = 1.0 / math.sqrt(self.dk)
        head = lambda t, part, h: t[..., (part * H + h) * dkp:(part * H + h + 1) * dkp]
        dx1, dbr1 = ops.layernorm_bwd(dy1, s.sv1, Lf.ln1_g, lens, g.ln1_g, g.ln1_b, dbias=g.o_b, want_drop=True,
                                      drop_pre_p=s.d1[0], seed_pre=s.d1[1], seed_dev=sd)
        ops.conv_wgrad(ops.pad_cols(dbr1, Cp)[..., :C], s.vec, (0,), out=g.o_w, accumulate=True)
        dvec = ops.conv_dgrad(dbr1, L.o_w, round_out=True)                                 # [B, T, H dkp]
        dqkv = torch.empty_like(s.qkv)
        dP = torch.empty(H, B, T, Tp, device=dy1.device, dtype=torch.float32)
        dPB = torch.empty(H, B, T, REL_COLS, device=dy1.device, dtype=torch.float32)
        Pd4, PB4 = s.Pd.view(H, B, T, Tp), s.PB.view(H, B, T, REL_COLS)
        for h in range(H):
            dv_h = dvec[..., h * dkp:(h + 1) * dkp]
            ops.bmm_nt(dv_h, head(s.qkv, 2, h), out=dP[h][..., :T])                       # dP = dO.v^T
            ops.conv_fwd(dv_h, L.ev, out=dPB[h])                                           # band part: dO.E_v^T
            ops.bmm_tn(Pd4[h][..., :T], dv_h, out=head(dqkv, 2, h), round_out=True)        # dv = P^T dO
            ops.conv_wgrad(PB4[h], dv_h, (0,), out=g.ev, accumulate=True)                  # dE_v = P_band^T dO
        ops.rel_band_add_(dP.view(H * B, T, Tp), dPB.view(H * B, T, REL_COLS), T, Wd)
        ops.softmax_bwd_(s.P, dP.view(H * B, T, Tp), T, alpha, s.att[0], s.att[1], sd)     # dP <- dS / sqrt(d_k)
        dR = ops.rel_band_gather(dP.view(H * B, T, Tp), T, Wd, REL_COLS).view(H, B, T, REL_COLS)
        for h in range(H):
            q_h, k_h = head(s.qkv, 0, h), head(s.qkv, 1, h)
            dq = ops.bmm_nn(dP[h][..., :T], k_h)                                           # dS.k
            ops.conv_dgrad(dR[h], L.ek, out=head(dqkv, 0, h), residual=dq, round_out=True)  # + dS_band.E_k
            ops.bmm_tn(dP[h][..., :T], q_h, out=head(dqkv, 1, h), round_out=True)          # dk = dS^T q
            ops.conv_wgrad(dR[h], q_h, (0,), out=g.ek, accumulate=True)                    # dE_k = dS_band^T q
        del dP, dPB, dR, dvec
        ops.conv_wgrad(dqkv, ops.pad_cols(s.x, Cp)[..., :C], (0,), out=g.qkv_w, accumulate=True)
        ops.colsum_(B * T, dqkv.shape[2], dqkv.shape[2], dqkv, g.qkv_b)
        return ops.conv_dgrad(dqkv, L.qkv_w, residual=dx1)


class TextEncoder(_RelTransformer):
    """Drop-in for python/xvapitch/model.py:1089 ``TextEncoder`` (see the module docstring): ``forward(tokens, x_lengths,
    lang_emb)`` -> (x [B, C, T], x_emb [B, T, hidden], x_mask [B, 1, T]); ``forward(x, x_lengths, stats=True, x_mask=...)`` ->
    (m_p, logs_p). Engine-facing entry points on channels-last tensors: ``forward_cl`` / ``stats_cl`` /
    ``stats_backward_cl`` / ``backward_cl``. ``flat`` is the one parameter tensor an optimizer sees."""

    def __init__(self, n_vocab, out_channels, hidden_channels, hidden_channels_ffn, num_heads, num_layers, kernel_size,
                 dropout_p, language_emb_dim=None, device=None, seed=1234):
        super().__init__()
        self.n_vocab, self.out_channels, self.hidden_channels = int(n_vocab), int(out_channels), int(hidden_channels)
        self.num_layers = int(num_layers)
        self.lang_dim = int(language_emb_dim or 0)
        if (2 * self.out_channels) % 32:
            raise NotImplementedError("2 * out_channels must be a multiple of 32")
        self._setup(self.hidden_channels + self.lang_dim, num_heads, hidden_channels_ffn, kernel_size, dropout_p)
        dev = _need_cuda(device)
        spec = [("emb", (self.n_vocab, self.hidden_channels))]
        for i in range(self.num_layers):
            spec += self._layer_spec(i)
        spec += [("proj_w", (2 * self.out_channels, self.Cp)), ("proj_b", (2 * self.out_channels,))]
        self._alloc(spec, dev, seed)
        self._stats_ctx = None
        self.reset_parameters(seed)

    # ------------------------------------------------------------------------------------------ parameters
    def _ref_spec(self):
        """(key, shape) of the reference module's state_dict, in its order (model.py:1117-1141, glow_tts.py:131-146,
        349-351, 420-447)."""
        C, Ce, F, k, dk = self.C, self.hidden_channels, self.hidden_channels_ffn, self.kernel_size, self.dk
        spec = [("emb.weight", (self.n_vocab, Ce))]
        for i in range(self.num_layers):
            a = f"encoder.attn_layers.{i}"
            spec += [(f"{a}.emb_rel_k", (1, 2 * self.W + 1, dk)), (f"{a}.emb_rel_v", (1, 2 * self.W + 1, dk))]
            for n in ("q", "k", "v", "o"):
                spec += [(f"{a}.conv_{n}.weight", (C, C, 1)), (f"{a}.conv_{n}.bias", (C,))]
        for i in range(self.num_layers):
            spec += [(f"encoder.norm_layers_1.{i}.gamma", (C,)), (f"encoder.norm_layers_1.{i}.beta", (C,))]
        for i in range(self.num_layers):
            f = f"encoder.ffn_layers.{i}"
            spec += [(f"{f}.conv_1.weight", (F, C, k)), (f"{f}.conv_1.bias", (F,)), (f"{f}.conv_2.weight", (C, F, k)),
                     (f"{f}.conv_2.bias", (C,))]
        for i in range(self.num_layers):
            spec += [(f"encoder.norm_layers_2.{i}.gamma", (C,)), (f"encoder.norm_layers_2.{i}.beta", (C,))]
        spec += [("proj.weight", (2 * self.out_channels, C, 1)), ("proj.bias", (2 * self.out_channels,))]
        return spec

    def _extra_to_ref(self, V, out, params):
        C = self.C
        out["emb.weight"] = V["emb"].clone()
        out["proj.weight"] = V["proj_w"][:, :C].reshape(2 * self.out_channels, C, 1).clone()
        out["proj.bias"] = V["proj_b"].clone()

    def _extra_from_ref(self, V, put, get):
        put(V["emb"], "emb.weight")
        put(V["proj_w"][:, :self.C], "proj.weight", lambda t: t[:, :, 0])
        put(V["proj_b"], "proj.bias")

    # ------------------------------------------------------------------------------------------ forward
    def forward_cl(self, tokens, lens, lang):
        """tokens int64 [B, T], lens int32 [B], lang [B, L] or None -> (x [B, T, C] channels-last with zero rows past the
        lengths, x_emb [B, T, hidden])."""
        B, T = tokens.shape
        Wr, Wf = self._weights()                                 # GEMM operands (rounded) / fp32 biases and LayerNorm
        keep = self.training
        self._site = 0
        scale = math.sqrt(self.hidden_channels)
        x, x_emb = ops.text_embed(tokens, Wf["emb"], lang, lens, scale, self.C)
        lens_rep = lens.repeat(self.num_heads)                   # one entry per (head, utterance): z = h * B + b
        saved = []
        for i in range(self.num_layers):
            L, Lf = self._layer(Wr, i), self._layer(Wf, i)
            y1, sa = self._attn_fwd(x, lens, lens_rep, L, Lf)
            y2, sf = self._ffn_fwd(y1, lens, L, Lf)
            if keep:
                sa.__dict__.update(sf.__dict__)
                saved.append(sa)
            x = y2
        self._ctx = _NS(saved=saved, tokens=tokens, lens=lens, lens_rep=lens_rep, lang=lang, T=T, B=B) if keep else None
        return x, x_emb

    def stats_cl(self, x, lens):
        """TextEncoder.forward(stats=True), model.py:1147-1150: x [B, T, C] -> stats [B, T, 2 out] = [m | logs], masked."""
        Wr, Wf = self._weights()
        stats = ops.conv_fwd(x, Wr["proj_w"][None, :, :self.C], bias=Wf["proj_b"], lens=lens)
        self._stats_ctx = (x, lens) if self.training else None
        return stats

    def forward(self, x, x_lengths, lang_emb=None, stats=False, x_mask=None, lang_emb_full=None):
        """The reference's signature and layouts (model.py:1143-1168). stats=False: x = tokens [B, T] -> (x [B, C, T],
        x_emb [B, T, hidden], x_mask [B, 1, T]). stats=True: x [B, C, T] -> (m, logs) [B, out, T]."""
        dev = self.flat.device
        if lang_emb_full is not None:
            raise NotImplementedError("lang_emb_full (a per-token language embedding) is not built")
        lens = torch.as_tensor(x_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
        if stats:
            xc = x.to(device=dev, dtype=torch.float32).transpose(1, 2).contiguous()
            st = self.stats_cl(xc, lens)
            o = self.out_channels
            return st[..., :o].transpose(1, 2), st[..., o:].transpose(1, 2)
        tokens = x.to(device=dev, dtype=torch.int64).contiguous()
        B, T = tokens.shape
        lang = None
        if self.lang_dim:
            if lang_emb is None:
                raise ValueError("this encoder was built with a language embedding: pass lang_emb [B, L, 1]")
            lang = lang_emb.to(device=dev, dtype=torch.float32).reshape(B, self.lang_dim).contiguous()
        xo, x_emb = self.forward_cl(tokens, lens, lang)
        mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(1)
        return xo.transpose(1, 2), x_emb, mask

    # ------------------------------------------------------------------------------------------ backward
    def stats_backward_cl(self, dstats):
        """dstats [B, T, 2 out] = dL/d[m | logs] (any values on the padded rows: they are masked here) -> dL/dx [B, T, C];
        accumulates the projection's gradients."""
        if self._stats_ctx is None:
            raise RuntimeError("stats_backward_cl() needs a stats call in training mode first")
        x, lens = self._stats_ctx
        B, T, C = x.shape
        Wr = self._views(self._w)
        G = self._views(self.flat.grad)
        # the forward multiplied by the mask (model.py:1148): zero the padded rows, and round -- d is a GEMM operand
        mask = (torch.arange(T, device=x.device)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(-1)
        d = (dstats.to(torch.float32) * mask).contiguous()
        ops.round_tf32_(d.view(-1), d.view(-1))
        ops.colsum_(B * T, d.shape[2], d.shape[2], d, G["proj_b"])
        xp = ops.pad_cols(x, self.Cp)
        ops.conv_wgrad(d, xp[..., :C], (0,), out=G["proj_w"][None, :, :C], accumulate=True)
        self._stats_ctx = None
        return ops.conv_dgrad(d, Wr["proj_w"][None, :, :C])

    def backward_cl(self, dx, dx_emb=None):
        """dx [B, T, C] = dL/dx of forward_cl's first output (rows past the lengths are ignored), dx_emb [B, T, hidden]
        (optional) = dL/dx_emb. Accumulates every parameter gradient into ``flat.grad``; returns dL/d(lang) [B, L] or None."""
        c = self._ctx
        if c is None:
            raise RuntimeError("backward_cl() needs a forward in training mode first")
        B, lens = c.B, c.lens
        Wr, Wf = self._views(self._w), self._views(self.flat.data)
        G = self._views(self.flat.grad)
        dy = dx.to(torch.float32).contiguous()
        for i in reversed(range(self.num_layers)):
            s = c.saved[i]
            L, Lf, g = self._layer(Wr, i), self._layer(Wf, i), self._layer(G, i)
            dy1 = self._ffn_bwd(dy, s, lens, L, Lf, g)
            dy = self._attn_bwd(dy1, s, lens, L, Lf, g)
        # ---- embedding and language embedding (model.py:1152-1165)
        scale = math.sqrt(self.hidden_channels)
        ops.text_embed_bwd_(c.tokens, dy, lens, self.hidden_channels, scale, G["emb"])
        if dx_emb is not None:
            ops.text_embed_bwd_(c.tokens, dx_emb.to(torch.float32).contiguous(), None, self.hidden_channels, scale, G["emb"])
        dlang = None
        if self.lang_dim:
            dlang = torch.zeros(B, self.lang_dim, device=dy.device, dtype=torch.float32)
            ops.colsum_items_(dy[..., self.hidden_channels:], dlang)
        self._ctx = None
        return dlang

    def backward(self, dx, dx_emb=None):
        """dx [B, C, T] (the reference's layout) -> dL/d(lang_emb) [B, L, 1] or None."""
        dlang = self.backward_cl(dx.transpose(1, 2), dx_emb)
        return None if dlang is None else dlang.unsqueeze(-1)


class RelativePositioningPitchEnergyEncoder(_RelTransformer):
    """Drop-in for python/xvapitch/model.py:1268 ``RelativePositioningPitchEnergyEncoder`` as xVAPitch builds its pitch
    predictor (model.py:154-168: out_channels = 1, 3 layers, the text encoder's channels + the 512-channel speaker embedding
    = 708 / 780): same constructor arguments, ``forward(x [B, T, hidden], x_lengths, speaker_emb [B, cond, 1])`` ->
    pitch_pred [B, 1, T], same state_dict keys and shapes.

    With out_channels = 1 the reference's last layer runs its FFN and throws the result away (glow_tts.py:479-483:
    ``x = self.proj(x)`` stands where ``norm_layers_2(x + y)`` would): here that FFN is not evaluated, and its six
    tensors (ffn_layers[-1].*, norm_layers_2[-1].*) are kept outside the optimizer's flat tensor, as plain state --
    in the reference they never receive a gradient, so AdamW never touches them, weight decay included.

    The 1-channel projection and its backward are the N = 1 / M = 1 tap-GEMM shapes of the HiFi-GAN generator's conv_post;
    LayerNorm over 708 / 780 channels is the 1024-channel instantiation of the LayerNorm kernels."""

    def __init__(self, out_channels, hidden_channels, hidden_channels_ffn, num_heads, num_layers, kernel_size, dropout_p,
                 conditioning_emb_dim=None, device=None, seed=1234):
        super().__init__()
        if int(out_channels) != 1:
            raise NotImplementedError("xVAPitch builds this module with out_channels = 1 (model.py:154-168)")
        self.out_channels, self.hidden_channels, self.num_layers = 1, int(hidden_channels), int(num_layers)
        self.cond_dim = int(conditioning_emb_dim or 0)
        self._setup(self.hidden_channels + self.cond_dim, num_heads, hidden_channels_ffn, kernel_size, dropout_p)
        dev = _need_cuda(device)
        spec = []
        for i in range(self.num_layers):
            spec += self._layer_spec(i, ffn=(i + 1 < self.num_layers))
        spec += [("proj_w", (1, self.Cp)), ("proj_b", (1,))]
        self._alloc(spec, dev, seed)
        self._dead = OrderedDict()                       # reference key -> tensor: parameters that are not on the path
        self.reset_parameters(seed)

    def _ref_spec(self):
        """(key, shape) of the reference module's state_dict in its order: ModuleLists first, ``encoder.proj`` (registered
        inside the constructor's loop, glow_tts.py:425-426) last."""
        C, F, k, dk, L = self.C, self.hidden_channels_ffn, self.kernel_size, self.dk, self.num_layers
        spec = []
        for i in range(L):
            a = f"encoder.attn_layers.{i}"
            spec += [(f"{a}.emb_rel_k", (1, 2 * self.W + 1, dk)), (f"{a}.emb_rel_v", (1, 2 * self.W + 1, dk))]
            for n in ("q", "k", "v", "o"):
                spec += [(f"{a}.conv_{n}.weight", (C, C, 1)), (f"{a}.conv_{n}.bias", (C,))]
        for i in range(L):
            spec += [(f"encoder.norm_layers_1.{i}.gamma", (C,)), (f"encoder.norm_layers_1.{i}.beta", (C,))]
        for i in range(L):
            f, o = f"encoder.ffn_layers.{i}", (C if i + 1 < L else 1)
            spec += [(f"{f}.conv_1.weight", (F, C, k)), (f"{f}.conv_1.bias", (F,)), (f"{f}.conv_2.weight", (o, F, k)),
                     (f"{f}.conv_2.bias", (o,))]
        for i in range(L):
            o = C if i + 1 < L else 1
            spec += [(f"encoder.norm_layers_2.{i}.gamma", (o,)), (f"encoder.norm_layers_2.{i}.beta", (o,))]
        spec += [("encoder.proj.weight", (1, C, 1)), ("encoder.proj.bias", (1,))]
        return spec

    def dead_keys(self):
        """The six reference parameters that never receive a gradient."""
        i = self.num_layers - 1
        return [f"encoder.ffn_layers.{i}.conv_1.weight", f"encoder.ffn_layers.{i}.conv_1.bias", f"encoder.ffn_layers.{i}.conv_2.weight",
                f"encoder.ffn_layers.{i}.conv_2.bias", f"encoder.norm_layers_2.{i}.gamma", f"encoder.norm_layers_2.{i}.beta"]

    def _extra_to_ref(self, V, out, params):
        out["encoder.proj.weight"] = V["proj_w"][:, :self.C].reshape(1, self.C, 1).clone()
        out["encoder.proj.bias"] = V["proj_b"].clone()
        if params:
            for k, t in self._dead.items():
                out[k] = t.clone()

    def _extra_from_ref(self, V, put, get):
        put(V["proj_w"][:, :self.C], "encoder.proj.weight", lambda t: t[:, :, 0])
        put(V["proj_b"], "encoder.proj.bias")
        for k in self.dead_keys():
            t = get(k)
            if t is not None:
                self._dead[k] = t.clone()

    # ------------------------------------------------------------------------------------------ forward / backward
    def forward_cl(self, x, lens, spk):
        """x [B, T, hidden] (the text encoder's output, zero rows past the lengths), lens int32 [B], spk [B, cond] or None
        -> pitch_pred [B, T, 1], zero rows past the lengths."""
        B, T, _ = x.shape
        C, hid = self.C, self.hidden_channels
        Wr, Wf = self._weights()
        keep = self.training
        self._site = 0
        mask = (torch.arange(T, device=x.device)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(-1)
        xin = torch.empty(B, T, C, device=x.device, dtype=torch.float32)           # cat(x, speaker_emb) * mask, model.py:1325-1346
        xin[..., :hid].copy_(x)
        if self.cond_dim:
            xin[..., hid:].copy_(spk.to(torch.float32).reshape(B, 1, self.cond_dim).expand(B, T, self.cond_dim))
        xin.mul_(mask)
        ops.round_tf32_(xin.view(-1), xin.view(-1))                                 # it is the first GEMM's operand
        lens_rep = lens.repeat(self.num_heads)
        saved, h = [], xin
        for i in range(self.num_layers):
            L, Lf = self._layer(Wr, i), self._layer(Wf, i)
            y1, sa = self._attn_fwd(h, lens, lens_rep, L, Lf)
            if i + 1 < self.num_layers:
                h, sf = self._ffn_fwd(y1, lens, L, Lf)
                sa.__dict__.update(sf.__dict__)
            else:
                h = y1
            if keep:
                saved.append(sa)
        pred = ops.conv_fwd(h, Wr["proj_w"][None, :, :C], bias=Wf["proj_b"], lens=lens)       # [B, T, 1]
        self._ctx = _NS(saved=saved, lens=lens, y_last=h, B=B, T=T, mask=mask) if keep else None
        return pred

    def forward(self, x, x_lengths=None, speaker_emb=None, stats=False, x_mask=None):
        """The reference's signature (model.py:1310): x [B, T, hidden], speaker_emb [B, cond, 1] -> [B, 1, T]."""
        dev = self.flat.device
        B, T, _ = x.shape
        lens = torch.as_tensor(x_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
        spk = None
        if self.cond_dim:
            if speaker_emb is None:
                raise ValueError("this predictor was built with a conditioning embedding: pass speaker_emb [B, cond, 1]")
            spk = speaker_emb.to(device=dev, dtype=torch.float32).reshape(B, self.cond_dim).contiguous()
        pred = self.forward_cl(x.to(device=dev, dtype=torch.float32).contiguous(), lens, spk)
        return pred.transpose(1, 2)

    def backward_cl(self, dpred, need_input_grad=False):
        """dpred [B, T] or [B, T, 1] = dL/d(pitch_pred) (rows past the lengths are ignored). Accumulates every parameter
        gradient into ``flat.grad``. need_input_grad: returns (dL/dx [B, T, hidden], dL/d(speaker_emb) [B, cond]) -- the
        reference detaches x (model.py:835) and the speaker embedding has no parameters, so training does not need them."""
        c = self._ctx
        if c is None:
            raise RuntimeError("backward_cl() needs a forward in training mode first")
        B, T, lens, C, Cp = c.B, c.T, c.lens, self.C, self.Cp
        Wr, Wf = self._views(self._w), self._views(self.flat.data)
        G = self._views(self.flat.grad)
        # the forward's final `* x_mask` (glow_tts.py:484); the single channel padded to one 32-column operand chunk
        d = (dpred.to(torch.float32).reshape(B, T, 1) * c.mask).contiguous()
        ops.round_tf32_(d.view(-1), d.view(-1))
        d32 = ops.pad_cols(d, 32)
        d1 = d32[..., :1]
        ops.colsum_(B * T, 1, 32, d32, G["proj_b"])
        ops.conv_wgrad(d1, ops.pad_cols(c.y_last, Cp)[..., :C], (0,), out=G["proj_w"][None, :, :C], accumulate=True)
        dy1 = ops.conv_dgrad(d1, Wr["proj_w"][None, :, :C])
        for i in reversed(range(self.num_layers)):
            s = c.saved[i]
            L, Lf, g = self._layer(Wr, i), self._layer(Wf, i), self._layer(G, i)
            if i + 1 < self.num_layers:
                dy1 = self._ffn_bwd(dy1, s, lens, L, Lf, g)
            dy1 = self._attn_bwd(dy1, s, lens, L, Lf, g)
        self._ctx = None
        if not need_input_grad:
            return None
        dspk = None
        if self.cond_dim:
            dspk = torch.zeros(B, self.cond_dim, device=dy1.device, dtype=torch.float32)
            ops.colsum_items_(dy1[..., self.hidden_channels:], dspk)
        return dy1[..., :self.hidden_channels], dspk

    def backward(self, dpred, need_input_grad=False):
        """dpred [B, 1, T] (the reference's layout)."""
        return self.backward_cl(dpred.transpose(1, 2), need_input_grad)
