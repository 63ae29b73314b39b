"""B200-native HiFi-GAN v1 training path: the reference's ``Generator`` (python/hifigan/models.py:81-137) with every
tensor operation issued through the C ABI of libxva_b200.so.

Same constructor argument (the config ``h``), same ``state_dict`` keys and shapes as the reference (``weight_g`` /
``weight_v`` / ``bias`` of every weight-normed conv: 234 keys), same ``forward(x)`` signature ([B, 80, T] mel ->
[B, 1, 256 T] waveform). Training adds ``backward(dy)``: there is no autograd graph over activations.

Design:
  * activations are channels-last [B, T, C] fp32; every conv is the tcgen05 tap-GEMM (k taps = k shifted TMA loads,
    dilation = shift stride, zero padding = TMA out-of-bounds fill);
  * each ResBlock1 step `x = c2(lrelu(c1(lrelu(x)))) + x` (models.py:41-48) is two launches: conv1 with the leaky ReLU
    in its epilogue, conv2 with the residual add in its epilogue and a second output leaky_relu(x) that is the next
    conv's operand -- no standalone activation or add kernels;
  * ConvTranspose1d(k = 2u, stride u) is two 2-tap GEMMs (output phases below / above u/2) writing column slices of
    the [B, T, u*Cout] view of the [B, u*T, Cout] output: no zero insertion, no col2im;
  * the weight-norm reparametrisation (w = g v / ||v||) and the re-packing of w into the kernel layout stay in PyTorch
    (tiny tensors; autograd carries dL/dw back to weight_g / weight_v), as SURVEY.md section 7 recommends.
"""
import math
import os

import torch
from torch import nn

from . import capi, ops

LRELU_SLOPE = 0.1


def _auto_streams(env, width):
    """Stream-level parallelism switch shared by _Side / _Branches / fastpitch.FastPitch: XVA_*_STREAMS = "0" off, a number
    = forced width, unset / "auto" = ``width`` while a CUDA graph is being captured and 0 otherwise. Measured on B200
    (profiles/r02_streams_ab.txt): inside a replayed graph the extra streams cost nothing on the host and win 2.6 %
    (FastPitch) / 15.6 % (HiFi-GAN); launched eagerly the stream switches and event records make the host-bound HiFi-GAN
    step 18 % slower, hence "auto"."""
    v = os.environ.get(env, "auto").strip().lower()
    if v in ("", "auto"):
        return width if ops.capturing() else 0
    try:
        return max(0, int(v))
    except ValueError:
        return 0


class _Side:
    """Weight / bias gradient launches go to a second stream (XVA_BWD_STREAMS: auto = inside graph capture, 0, 1). They only read tensors that already exist and accumulate into the packed gradient
    arena / the bias .grad tensors, allocate nothing, and nothing on the main stream depends on them until the arena is
    unpacked -- so the small-channel weight gradients (52 CTAs on 148 SMs, 160 us each at B = 16 x 8192) overlap the
    input-gradient chain instead of serialising with it. Inputs are record_stream()-ed: their memory is not reused
    before the side stream is done with it; join() runs where the gradients are consumed (_WnPacker.unpack_grads)."""
    enabled = None          # None: by XVA_BWD_STREAMS (default auto); True / False: forced (tests)
    streams = {}            # launching stream (handle) -> its side stream
    used = []
    stream = None           # the side stream of the most recent run() (tests look at it)

    @classmethod
    def on(cls):
        return bool(_auto_streams("XVA_BWD_STREAMS", 1)) if cls.enabled is None else bool(cls.enabled)

    @classmethod
    def run(cls, fn, *inputs):
        """One side stream PER LAUNCHING STREAM: the weight gradients of the three ResBlock branches of a generator stage
        (or of the eight sub-discriminators) do not queue behind each other on a single stream -- a weight-gradient
        launch fills one to two waves of CTAs, and a step issues ~10 ms of them (gemm_tc_kernel<0, 1> in
        profiles/r02_timeline_hifigan_after.txt). Launches from one branch stay ordered among themselves (two passes of a spectral-normed sub-discriminator
        accumulate into the same bias gradients)."""
        if not cls.on():
            return fn()
        cur = torch.cuda.current_stream()
        side = cls.streams.get(cur.cuda_stream)
        if side is None:
            side = cls.streams[cur.cuda_stream] = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
        for t in inputs:
            t.record_stream(side)
        if side not in cls.used:
            cls.used.append(side)
        cls.stream = side

    @classmethod
    def join(cls):
        if cls.used:
            cur = torch.cuda.current_stream()
            for side in cls.used:
                cur.wait_stream(side)
            cls.used = []


class _Branches:
    """The sub-discriminators of MPD / MSD are independent chains of small launches; with n > 0 sub-discriminator i runs
    -- forward, loss gradients and backward -- on stream i mod n (XVA_DISC_STREAMS: auto = 8 inside graph capture, 0, n;
    XVA_GEN_STREAMS the same for the three ResBlocks of an MRF stage). Discipline that keeps the caching allocator safe without
    record_stream: a branch starts by waiting for the main stream (fork) and everything in it, torch ops included, runs on
    its stream; the main stream waits for every branch (join) before it touches what they produced; a tensor created in
    a branch is freed in it or after the join; a main-stream tensor a branch reads is kept referenced until the join."""
    n = None                # None: by XVA_DISC_STREAMS (default auto); an int: forced (tests)
    n_gen = None            # None: by XVA_GEN_STREAMS; an int: forced
    streams = []
    open_streams = []

    class _Ctx:
        def __init__(self, stream):
            self.stream = stream

        def __enter__(self):
            self.stream.wait_stream(torch.cuda.current_stream())
            self.cm = torch.cuda.stream(self.stream)
            self.cm.__enter__()

        def __exit__(self, *exc):
            self.cm.__exit__(*exc)
            if self.stream not in _Branches.open_streams:
                _Branches.open_streams.append(self.stream)
            return False

    class _Null:
        def __enter__(self):
            return None

        def __exit__(self, *exc):
            return False

    @classmethod
    def width(cls):
        return _auto_streams("XVA_DISC_STREAMS", 8) if cls.n is None else int(cls.n)

    @classmethod
    def gen_width(cls):
        return (3 if _auto_streams("XVA_GEN_STREAMS", 1) else 0) if cls.n_gen is None else int(cls.n_gen)

    @classmethod
    def on(cls):
        return cls.width() > 0

    @classmethod
    def branch(cls, i, n=None):
        n = cls.width() if n is None else n
        if n <= 0:
            return cls._Null()
        while len(cls.streams) < n:
            cls.streams.append(torch.cuda.Stream())
        return cls._Ctx(cls.streams[i % n])

    @classmethod
    def join(cls):
        if not cls.open_streams:
            return
        main = torch.cuda.current_stream()
        for st in cls.open_streams:
            main.wait_stream(st)
        cls.open_streams = []


class _WNConv(nn.Module):
    """Parameters of one weight-normed Conv1d / ConvTranspose1d with the reference's names (bias, weight_g, weight_v)."""

    def __init__(self, cin, cout, k, dilation=1, transposed=False, stride=1):
        super().__init__()
        self.cin, self.cout, self.k, self.dilation, self.transposed, self.stride = cin, cout, k, dilation, transposed, stride
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        self.bias = nn.Parameter(torch.zeros(cout))
        self.weight_g = nn.Parameter(torch.ones(shape[0], 1, 1))
        self.weight_v = nn.Parameter(torch.zeros(shape))

    def weight(self):
        v = self.weight_v
        return v * (self.weight_g / v.flatten(1).norm(dim=1).view(-1, 1, 1))

    @property
    def shifts(self):
        half = (self.k - 1) // 2
        return tuple((j - half) * self.dilation for j in range(self.k))


class _PlainConv(nn.Module):
    """Parameters of one Conv1d without weight norm, with torch's names (weight, bias): what remove_weight_norm leaves of
    conv_pre / conv_post in the xVAPitch waveform decoder, and its cond_layer (python/xvapitch/hifigan.py:222-232)."""

    transposed, stride = False, 1

    def __init__(self, cin, cout, k, dilation=1, bias=True, bias_first=True):
        """bias_first: state_dict order of a conv whose weight norm was removed (bias, weight -- remove_weight_norm
        re-registers the weight last); a conv that never had one has torch's (weight, bias)."""
        super().__init__()
        self.cin, self.cout, self.k, self.dilation = cin, cout, k, dilation
        if not bias_first:
            self.weight = nn.Parameter(torch.zeros(cout, cin, k))
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))
        else:
            self.register_parameter("bias", None)
        if bias_first:
            self.weight = nn.Parameter(torch.zeros(cout, cin, k))

    shifts = _WNConv.shifts


def _wv(m):
    """The parameter whose dim-0 rows the packer reads: weight_v, or the weight itself of a plain convolution."""
    return m.weight if isinstance(m, _PlainConv) else m.weight_v


def _eff_weight(m):
    return m.weight if isinstance(m, _PlainConv) else m.weight()


def _bias(m):
    return None if m.bias is None else m.bias.detach()


class _WnPacker:
    """Effective weights w = g * v / ||v|| of a set of weight-normed convolutions, written by ONE kernel launch
    (xva_wn_pack_fwd) into a flat arena in the layouts the tap-GEMM reads, and the way back (xva_wn_pack_bwd: gradient
    arena -> .grad of weight_g / weight_v) in one more. Replaces ~20 PyTorch launches per convolution and step."""

    def __init__(self):
        self.items = []          # (module, flags, ld, og, f, cg, tap_off list (arena-relative))
        self.layout = {}         # key -> [(offset, shape)]
        self.size = 0
        self.table = None
        self._ptrs = None

    def _alloc(self, key, shape):
        off = self.size
        n = int(math.prod(shape))
        self.size += (n + 63) // 64 * 64          # 256-byte alignment (TMA needs 16)
        self.layout.setdefault(key, []).append((off, tuple(shape)))
        return off

    def add_conv(self, key, m, cout, cg, k, order=None, og=None, f=1, ld=None):
        """Conv weight v [cout, cg, k] -> [k (in `order`), cout, f * cg] with group r's columns at ((r / og) % f) * cg.
        ld > f * cg: rows padded with zero columns (an input width that is not a multiple of 32)."""
        ld = f * cg if ld is None else int(ld)
        assert ld >= f * cg
        off = self._alloc(key, (k, cout, ld))
        order = list(range(k)) if order is None else list(order)
        taps = [0] * k
        for pos, j in enumerate(order):
            taps[j] = off + pos * cout * ld
        self.items.append((m, capi.WN_PLAIN if isinstance(m, _PlainConv) else 0, ld, og if og else cout, f, cg, taps))

    def add_flat(self, key, m, cout, k):
        """First discriminator layer (one input channel): [cout, k], not rounded (it runs on CUDA cores in fp32)."""
        off = self._alloc(key, (cout, k))
        self.items.append((m, capi.WN_NO_ROUND, k, cout, 1, 0, [off + j for j in range(k)]))

    def add_transposed(self, key, m, cin, cout, u):
        """ConvTranspose1d(k = 2u, stride u) weight v [cin, cout, 2u] -> the two 2-tap phase groups of Generator.forward."""
        p = u // 2
        lo = self._alloc(key, (2, (u - p) * cout, cin))
        hi = self._alloc(key, (2, p * cout, cin))
        mat = cout * cin
        taps = [0] * (2 * u)
        for kk in range(2 * u):
            if p <= kk < u:
                taps[kk] = lo + (kk - p) * mat
            elif kk >= p + u:
                taps[kk] = lo + (u - p) * mat + (kk - p - u) * mat
            elif kk < p:
                taps[kk] = hi + kk * mat
            else:
                taps[kk] = hi + p * mat + (kk - u) * mat
        self.items.append((m, capi.WN_TRANSPOSED, cin, 1, 1, 0, taps))

    def finalize(self, device):
        self.arena = torch.zeros(self.size, device=device, dtype=torch.float32)    # off-diagonal blocks stay zero
        self.garena = torch.zeros(self.size, device=device, dtype=torch.float32)
        view = lambda a: {k: tuple(a[o:o + int(math.prod(sh))].view(sh) for o, sh in v) for k, v in self.layout.items()}
        self.W, self.gW = view(self.arena), view(self.garena)
        self.rows = sum(_wv(m).shape[0] for m, *_ in self.items)
        self.max_inner = max(_wv(m).numel() // _wv(m).shape[0] for m, *_ in self.items)

    def _sync(self, grads):
        ptrs = []
        for m, *_ in self.items:
            v, g = _wv(m), getattr(m, "weight_g", None)        # g is None: a plain convolution (XVA_WN_PLAIN)
            if grads:
                for prm in (v, g):
                    if prm is not None and prm.grad is None:
                        prm.grad = torch.zeros_like(prm)
            ptrs.append((v.data_ptr(), g.data_ptr() if g is not None else 0,
                         v.grad.data_ptr() if v.grad is not None else 0,
                         g.grad.data_ptr() if g is not None and g.grad is not None else 0))
        if ptrs == self._ptrs:
            return
        import ctypes as C
        arr = (capi.WnDesc * len(self.items))()
        row = 0
        for d, (m, flags, ld, og, f, cg, taps), (pv, pg, pdv, pdg) in zip(arr, self.items, ptrs):
            v = _wv(m)
            assert v.is_contiguous() and (pg == 0 or m.weight_g.is_contiguous())
            d.v, d.g, d.dv, d.dg = pv, pg, pdv, pdg
            d.dst, d.ddst = self.arena.data_ptr(), self.garena.data_ptr()
            d.rows, d.inner, d.k, d.flags = v.shape[0], v.numel() // v.shape[0], len(taps), flags
            d.ld, d.og, d.f, d.cg = ld, og, f, cg
            d.row_start = row
            row += v.shape[0]
            for j, t in enumerate(taps):
                d.tap_off[j] = t
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = host.to(self.arena.device)
        self._ptrs = ptrs

    def pack(self):
        self._sync(False)
        capi.call("xva_wn_pack_fwd", ops._p(self.table), len(self.items), self.rows, self.max_inner, ops._stream())
        return self.W

    def zero_grads(self):
        self.garena.zero_()
        return self.gW

    def unpack_grads(self):
        _Side.join()
        self._sync(True)
        capi.call("xva_wn_pack_bwd", ops._p(self.table), len(self.items), self.rows, self.max_inner, ops._stream())


class _SnPacker:
    """Spectral-normed convolutions of a model (torch.nn.utils.spectral_norm: weight_orig, weight_u, weight_v): power
    iteration, sigma, w = weight_orig / sigma and the re-packing for the tap-GEMM in 5 launches per call
    (xva_sn_pack_fwd; 3 in eval mode), the way back to weight_orig.grad in 3 (xva_sn_pack_bwd). ``slots`` calls per
    forward keep their own packed weights, gradient arena and (u, v, sigma): the reference runs the discriminator on the
    real and on the generated waveform separately, each call advancing the power iteration (models.py:251-252)."""

    def __init__(self, slots=2):
        self.slots = slots
        self.items = []
        self.layout = {}
        self.size = 0
        self._tables = [None] * slots
        self._ptrs = [None] * slots

    _alloc = _WnPacker._alloc

    def add_conv(self, key, m, cout, cg, k, order=None, og=None, f=1, ld=None):
        ld = f * cg if ld is None else int(ld)
        off = self._alloc(key, (k, cout, ld))
        order = list(range(k)) if order is None else list(order)
        taps = [0] * k
        for pos, j in enumerate(order):
            taps[j] = off + pos * cout * ld
        self.items.append((m, 0, ld, og if og else cout, f, cg, taps))

    def add_flat(self, key, m, cout, k):
        off = self._alloc(key, (cout, k))
        self.items.append((m, capi.WN_NO_ROUND, k, cout, 1, 0, [off + j for j in range(k)]))

    def finalize(self, device):
        z = lambda n: torch.zeros(n, device=device, dtype=torch.float32)
        self.arena = [z(self.size) for _ in range(self.slots)]
        self.garena = [z(self.size) for _ in range(self.slots)]
        view = lambda a: {k: tuple(a[o:o + int(math.prod(sh))].view(sh) for o, sh in v) for k, v in self.layout.items()}
        self.W, self.gW = [view(a) for a in self.arena], [view(a) for a in self.garena]
        self.rows = sum(m.weight_orig.shape[0] for m, *_ in self.items)
        geo = [(m.weight_orig.shape[0], m.weight_orig.numel() // m.weight_orig.shape[0]) for m, *_ in self.items]
        self.max_inner = max(i for _, i in geo)
        self.blocks = sum(((i + 255) // 256) * ((r + 63) // 64) for r, i in geo)
        self.work = [[z(((r + 63) // 64) * i + r + 2) for r, i in geo] for _ in range(self.slots)]
        self.u_sav = [[z(r) for r, _ in geo] for _ in range(self.slots)]
        self.v_sav = [[z(i) for _, i in geo] for _ in range(self.slots)]

    def _sync(self, slot, grads):
        ptrs = []
        for m, *_ in self.items:
            if grads and m.weight_orig.grad is None:
                m.weight_orig.grad = torch.zeros_like(m.weight_orig)
            ptrs.append((m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(),
                         m.weight_orig.grad.data_ptr() if m.weight_orig.grad is not None else 0))
        if ptrs == self._ptrs[slot]:
            return
        arr = (capi.SnDesc * len(self.items))()
        row = blk = 0
        for n, (d, (m, flags, ld, og, f, cg, taps), (pw, pu, pv, pdw)) in enumerate(zip(arr, self.items, ptrs)):
            w = m.weight_orig
            assert w.is_contiguous() and m.weight_u.is_contiguous() and m.weight_v.is_contiguous()
            rows, inner = w.shape[0], w.numel() // w.shape[0]
            d.w, d.u, d.v, d.dw = pw, pu, pv, pdw
            d.u_sav, d.v_sav = self.u_sav[slot][n].data_ptr(), self.v_sav[slot][n].data_ptr()
            d.dst, d.ddst, d.work = self.arena[slot].data_ptr(), self.garena[slot].data_ptr(), self.work[slot][n].data_ptr()
            d.rows, d.inner, d.k, d.flags = rows, inner, len(taps), flags
            d.ld, d.og, d.f, d.cg = ld, og, f, cg
            d.row_start, d.blk_start = row, blk
            row += rows
            blk += ((inner + 255) // 256) * ((rows + 63) // 64)
            for j, t in enumerate(taps):
                d.tap_off[j] = t
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._tables[slot] = host.to(self.arena[slot].device)
        self._ptrs[slot] = ptrs

    def pack(self, slot, training):
        self._sync(slot, False)
        capi.call("xva_sn_pack_fwd", ops._p(self._tables[slot]), len(self.items), self.rows, self.blocks, self.max_inner,
                  int(bool(training)), ops._stream())
        return self.W[slot]

    def zero_grads(self):
        for a in self.garena:
            a.zero_()

    def unpack_grads(self):
        _Side.join()
        for slot in range(self.slots):
            self._sync(slot, True)
            capi.call("xva_sn_pack_bwd", ops._p(self._tables[slot]), len(self.items), self.rows, self.blocks, self.max_inner,
                      ops._stream())


class ResBlock1(nn.Module):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([_WNConv(channels, channels, kernel_size, d) for d in dilation])
        self.convs2 = nn.ModuleList([_WNConv(channels, channels, kernel_size, 1) for _ in dilation])


class Generator(nn.Module):
    """Drop-in for hifigan/models.py:81 ``Generator(h)`` (config_v1: resblock '1')."""

    def __init__(self, h, device=None, seed=1234, in_channels=80, cond_channels=0, conv_pre_weight_norm=True,
                 conv_post_weight_norm=True, conv_post_bias=True):
        """The keyword arguments after `seed` are the xVAPitch waveform decoder's (HifiganGenerator below); their
        defaults are hifigan/models.py's generator."""
        super().__init__()
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        if str(h.resblock) != "1":
            raise NotImplementedError("only ResBlock1 (config_v1.json) is built")
        c0 = h.upsample_initial_channel
        self.in_channels = int(in_channels)
        self.in_cols = (self.in_channels + 31) // 32 * 32      # operand rows are whole 32-column chunks (zero tail)
        self.conv_pre = _WNConv(self.in_channels, c0, 7) if conv_pre_weight_norm else _PlainConv(self.in_channels, c0, 7)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            if k != 2 * u or u % 2:
                raise NotImplementedError(f"upsample kernel {k} / rate {u}: only k = 2u with even u is built")
            self.ups.append(_WNConv(c0 // 2 ** i, c0 // 2 ** (i + 1), k, transposed=True, stride=u))
        self.resblocks = nn.ModuleList()
        for i in range(len(self.ups)):
            ch = c0 // 2 ** (i + 1)
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(ResBlock1(h, ch, k, tuple(d)))
        self.conv_post = _WNConv(ch, 1, 7) if conv_post_weight_norm else _PlainConv(ch, 1, 7, bias=conv_post_bias)
        if conv_post_weight_norm and not conv_post_bias:
            raise NotImplementedError("a weight-normed conv_post without bias is not a configuration the reference uses")
        if cond_channels > 0:
            if cond_channels % 32:
                raise NotImplementedError(f"cond_channels={cond_channels}: must be a multiple of 32")
            self.cond_layer = _PlainConv(int(cond_channels), c0, 1, bias_first=False)
        self.reset_parameters(seed)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if dev.type != "cuda":
            raise capi.XvaError("Generator (B200 build) needs a CUDA device: there is no CPU path")
        capi.load()
        capi.call("xva_device_check", dev.index or 0)
        self.to(dev)
        self._ctx = None
        self._packer = None

    def _get_packer(self):
        if self._packer is None:
            pk = _WnPacker()
            for name, m in self.named_modules():
                if not isinstance(m, (_WNConv, _PlainConv)):
                    continue
                if m.transposed:
                    pk.add_transposed(name, m, m.cin, m.cout, m.stride)
                else:
                    pk.add_conv(name, m, m.cout, m.cin, m.k)
            pk.finalize(_wv(self.conv_pre).device)
            self._packer = pk
        return self._packer

    def reset_parameters(self, seed=1234):
        """Reference init: N(0, 0.01) weights for ups / resblocks / conv_post (utils.py:23-26), torch default for
        conv_pre, g = ||v|| (what weight_norm does at wrap time); generated on the CPU from one seed."""
        g = torch.Generator().manual_seed(int(seed))
        with torch.no_grad():
            for name, m in self.named_modules():
                if not isinstance(m, (_WNConv, _PlainConv)):
                    continue
                v = _wv(m)
                fan_in = v.shape[1] * v.shape[2]
                if name in ("conv_pre", "cond_layer") or isinstance(m, _PlainConv):
                    bound = 1.0 / math.sqrt(fan_in)
                    v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
                else:
                    v.copy_(torch.randn(v.shape, generator=g) * 0.01)
                if isinstance(m, _WNConv):
                    m.weight_g.copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1))
                if m.bias is not None:
                    m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) / math.sqrt(fan_in))

    # ------------------------------------------------------------------------------------------ weights
    def _pack(self):
        """Effective weights in the kernel layout [taps, Cout, Cin] (tf32-rounded by the GEMM's own operand rule: the
        packed copies are produced by torch here and rounded by xva_round_tf32), under autograd so that
        backward() can hand dL/d(packed) back to weight_g / weight_v."""
        packed = {}
        for name, m in self.named_modules():
            if not isinstance(m, (_WNConv, _PlainConv)):
                continue
            w = _eff_weight(m)
            if not m.transposed:
                packed[name] = (w.permute(2, 0, 1).contiguous(),)
            else:
                u, p = m.stride, m.stride // 2
                wk = w.permute(2, 1, 0)                                   # [k, Cout, Cin]
                cat = lambda a, b: torch.stack([a.reshape(-1, m.cin), b.reshape(-1, m.cin)]).contiguous()
                # phases r' < u - p: taps (shift 0, shift -1) = kernel columns (r'+p, r'+p+u)
                lo = cat(wk[p:u], wk[p + u:2 * u])
                # phases r' >= u - p: taps (shift +1, shift 0) = kernel columns (r'+p-u, r'+p)
                hi = cat(wk[0:p], wk[u:u + p])
                packed[name] = (lo, hi)
        return packed

    def _rounded(self, t):
        out = torch.empty_like(t)
        ops.round_tf32_(t.detach().reshape(-1), out.reshape(-1))
        return out

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x, cond_emb=None, g=None):
        """Generator.forward, models.py:110-128. x [B, 80, T] -> [B, 1, 256 T]. In training mode the tensors backward()
        needs are kept until the next forward(). g [B, cond_channels, 1]: the xVAPitch decoder's per-utterance
        conditioning, cond_layer(g) added to conv_pre's output before the first leaky ReLU
        (python/xvapitch/hifigan.py:248-250); ignored, as there, by a generator built without cond_channels."""
        if cond_emb is not None:
            raise NotImplementedError("USE_EMB_CONDITIONING is off in config_v1.json")
        B, cin, T = x.shape
        if cin != self.in_channels:
            raise ValueError(f"input has {cin} channels, the generator was built for {self.in_channels}")
        keep = self.training
        W = self._get_packer().pack()      # tf32-rounded effective weights of all 86 convolutions, one launch
        # [B, T, 80] channels-last operand in rows of 96 floats (zero tail): the weight-gradient GEMM reads it MN-major
        # in 32-column chunks
        melp = torch.zeros(B, T, self.in_cols, device=x.device, dtype=torch.float32)
        mel = melp[..., :cin]
        mel.copy_(x.to(torch.float32).transpose(1, 2))
        ops.round_tf32_(melp.reshape(-1), melp.reshape(-1))
        ctx = {"mel": mel, "stages": [], "W": W, "B": B, "cond": None}
        # conv_pre; only leaky_relu(conv_pre(x)) is ever read (models.py:111,115)
        if g is not None and hasattr(self, "cond_layer"):
            # c = cond_layer(g) is one [B, cond] x [cond, C0] product (a 1-tap convolution over the B conditioning
            # vectors taken as the rows of one item); conv_pre's epilogue adds it to every frame of its utterance
            # through the residual slot read with row stride 0, and writes leaky_relu(.) as its second output
            cl = self.cond_layer
            if g.shape[0] != B or g.shape[1] != cl.cin or g.shape[2] != 1:
                raise ValueError(f"g: expected [{B}, {cl.cin}, 1], got {tuple(g.shape)}")
            gr = self._rounded(g.to(torch.float32).reshape(1, B, cl.cin).contiguous())
            c = ops.conv_fwd(gr, W["cond_layer"][0], (0,), bias=cl.bias.detach())                     # [1, B, C0]
            pre = torch.empty(B, T, cl.cout, device=x.device, dtype=torch.float32)
            a = torch.empty_like(pre)
            ops.conv_fwd(mel, W["conv_pre"][0], self.conv_pre.shifts, out=pre, bias=_bias(self.conv_pre),
                         residual=c.view(B, 1, cl.cout).expand(B, T, cl.cout), out_act=a, out_act_slope=LRELU_SLOPE)
            ctx["cond"] = gr
        else:
            a = ops.conv_fwd(mel, W["conv_pre"][0], self.conv_pre.shifts, bias=_bias(self.conv_pre),
                             act_slope=LRELU_SLOPE, round_out=True)
        for i in range(self.num_upsamples):
            up = self.ups[i]
            u, p, cout = up.stride, up.stride // 2, up.cout
            Tin = a.shape[1]
            bias_rep = up.bias.detach().repeat(u)
            xu = torch.empty(B, Tin, u * cout, device=a.device, dtype=torch.float32)  # = [B, u*Tin, cout]
            au = torch.empty_like(xu)
            nlo = (u - p) * cout
            ops.conv_fwd(a, W[f"ups.{i}"][0], (0, -1), out=xu[..., :nlo], bias=bias_rep[:nlo], out_act=au[..., :nlo],
                         out_act_slope=LRELU_SLOPE)
            ops.conv_fwd(a, W[f"ups.{i}"][1], (1, 0), out=xu[..., nlo:], bias=bias_rep[nlo:], out_act=au[..., nlo:],
                         out_act_slope=LRELU_SLOPE)
            x0 = xu.view(B, Tin * u, cout)
            a0 = au.view(B, Tin * u, cout)
            stage = {"a_in": a, "a0": a0, "blocks": []}
            ys = []
            for j in range(self.num_kernels):
                rb = self.resblocks[i * self.num_kernels + j]
                name = f"resblocks.{i * self.num_kernels + j}"
                with _Branches.branch(j, _Branches.gen_width()):     # the ResBlocks of a stage only share their input
                    xr, ar = x0, a0
                    saved = []
                    for m in range(3):
                        c1, c2 = rb.convs1[m], rb.convs2[m]
                        t = ops.conv_fwd(ar, W[f"{name}.convs1.{m}"][0], c1.shifts, bias=c1.bias.detach(),
                                         act_slope=LRELU_SLOPE, round_out=True)
                        last = m == 2
                        xn = torch.empty_like(xr)
                        an = None if last else torch.empty_like(xr)
                        ops.conv_fwd(t, W[f"{name}.convs2.{m}"][0], c2.shifts, out=xn, bias=c2.bias.detach(), residual=xr,
                                     out_act=an, out_act_slope=LRELU_SLOPE)
                        saved.append((ar, t))
                        xr, ar = xn, an
                ys.append(xr)
                stage["blocks"].append(saved)
            _Branches.join()
            slope = LRELU_SLOPE if i + 1 < self.num_upsamples else 0.01               # models.py:115 vs :124
            a = ops.mean3_lrelu(ys[0], ys[1], ys[2], slope)
            ctx["stages"].append(stage)
        y = ops.conv_fwd(a, W["conv_post"][0], self.conv_post.shifts, bias=_bias(self.conv_post), tanh=True)
        ctx["a_last"], ctx["y"] = a, y
        self._ctx = ctx if keep else None
        return y.view(B, 1, -1)

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, dy, need_input_grad=False):
        """dy = dL/d(output) [B, 1, 256 T]. Accumulates .grad of every parameter (bias directly, weight_g / weight_v
        through the packer's weight-norm backward). need_input_grad: also return (dL/dx [B, in_channels, T], dL/dg
        [B, cond_channels, 1] or None) -- the xVAPitch decoder's input is the posterior encoder's latent, which trains."""
        ctx = self._ctx
        if ctx is None:
            raise RuntimeError("backward() needs a forward() in training mode first")
        B, W = ctx["B"], ctx["W"]
        gW = self._get_packer().zero_grads()      # packed-weight gradients: one arena, one fill
        mods = dict(self.named_modules())

        def bias_grad(name, d, cols, ld=None):
            b = mods[name].bias
            if b is None:
                return
            if b.grad is None:
                b.grad = torch.zeros_like(b)
            _Side.run(lambda: ops.colsum_(d.shape[0] * d.shape[1], cols, ld if ld is not None else d.shape[2], d, b.grad), d)

        def wgrad(dy_, x_, shifts, out):
            _Side.run(lambda: ops.conv_wgrad(dy_, x_, shifts, out=out, accumulate=True), dy_, x_)

        # conv_post + tanh
        y, a = ctx["y"], ctx["a_last"]
        Tl = a.shape[1]
        dpre = ops.tanh_bwd(dy.reshape(-1).to(torch.float32).contiguous(), y.reshape(-1), 32).view(B, Tl, 32)
        dp1 = dpre[..., :1]
        wgrad(dp1, a, self.conv_post.shifts, gW["conv_post"][0])
        bias_grad("conv_post", dpre, 1, 32)
        for i in reversed(range(self.num_upsamples)):
            stage = ctx["stages"][i]
            slope = LRELU_SLOPE if i + 1 < self.num_upsamples else 0.01
            # gradient wrt each ResBlock output y_j: (1/3) * lrelu'(mean) * d(a); a carries the sign of the mean
            if i + 1 == self.num_upsamples:
                dyj = ops.conv_dgrad(dp1, W["conv_post"][0], self.conv_post.shifts, gate=a, gate_slope=slope,
                                     alpha=1.0 / 3.0, round_out=True)
            else:
                dyj = self._ups_dgrad(i + 1, d_up, a, slope, W)
            a0 = stage["a0"]
            dx_blocks = []
            for j in range(self.num_kernels):
                rb = self.resblocks[i * self.num_kernels + j]
                name = f"resblocks.{i * self.num_kernels + j}"
                with _Branches.branch(j, _Branches.gen_width()):
                    G = dyj
                    for m in reversed(range(3)):
                        c1, c2 = rb.convs1[m], rb.convs2[m]
                        ar, t = stage["blocks"][j][m]
                        bias_grad(f"{name}.convs2.{m}", G, c2.cout)
                        wgrad(G, t, c2.shifts, gW[f"{name}.convs2.{m}"][0])
                        dt = ops.conv_dgrad(G, W[f"{name}.convs2.{m}"][0], c2.shifts, gate=t, gate_slope=LRELU_SLOPE,
                                            round_out=True)
                        bias_grad(f"{name}.convs1.{m}", dt, c1.cout)
                        wgrad(dt, ar, c1.shifts, gW[f"{name}.convs1.{m}"][0])
                        G = ops.conv_dgrad(dt, W[f"{name}.convs1.{m}"][0], c1.shifts, gate=ar, gate_slope=LRELU_SLOPE,
                                           residual=G, round_out=True)
                dx_blocks.append(G)
            _Branches.join()
            d_up = ops.sum3(dx_blocks[0], dx_blocks[1], dx_blocks[2])       # dL/d(ups[i] output), [B, u*Tin, cout]
            a = stage["a_in"]
            up = self.ups[i]
            u, p, cout = up.stride, up.stride // 2, up.cout
            Tin = a.shape[1]
            dv = d_up.view(B, Tin, u * cout)
            nlo = (u - p) * cout
            wgrad(dv[..., :nlo], a, (0, -1), gW[f"ups.{i}"][0])
            wgrad(dv[..., nlo:], a, (1, 0), gW[f"ups.{i}"][1])
            bias_grad(f"ups.{i}", d_up, cout)
        # ups[0] input is leaky_relu(conv_pre(x)): gate by its sign, then conv_pre's own gradients
        dpre0 = self._ups_dgrad(0, d_up, a, LRELU_SLOPE, W, alpha=1.0)
        wgrad(dpre0, ctx["mel"], self.conv_pre.shifts, gW["conv_pre"][0])
        bias_grad("conv_pre", dpre0, self.conv_pre.cout)
        dx = dg = None
        gr = ctx["cond"]
        if gr is not None:
            # d(cond_layer output)[b] = sum over the utterance's frames of dpre0: one column sum per item, then the
            # layer's own gradients as the 1-tap convolution over [1, B, cond] that the forward ran
            cl = self.cond_layer
            dc = torch.zeros(1, B, cl.cout, device=dpre0.device, dtype=torch.float32)
            ops.colsum_items_(dpre0, dc[0])
            bias_grad("cond_layer", dc, cl.cout)
            dcr = self._rounded(dc)
            wgrad(dcr, gr, (0,), gW["cond_layer"][0])
            if need_input_grad:
                dg = ops.conv_dgrad(dcr, W["cond_layer"][0], (0,)).view(B, cl.cin, 1)
        if need_input_grad:
            dx = ops.conv_dgrad(dpre0, W["conv_pre"][0], self.conv_pre.shifts)[..., :self.in_channels].transpose(1, 2)

        # packed-weight gradients -> weight_g / weight_v (.grad accumulated), one launch
        self._get_packer().unpack_grads()
        self._ctx = None
        if need_input_grad:
            return dx.contiguous(), dg

    def _ups_dgrad(self, i, d_up, a_in, slope, W, alpha=1.0 / 3.0):
        """Input gradient of ups[i] (two 2-tap dgrads, summed through the residual slot), times the derivative of the
        leaky ReLU that produced its input (gate = a_in) and the 1/3 of the MRF mean that precedes it."""
        up = self.ups[i]
        u, p, cout = up.stride, up.stride // 2, up.cout
        B, Tin, _ = a_in.shape
        dv = d_up.view(B, Tin, u * cout)
        nlo = (u - p) * cout
        # alpha and the gate are linear, so they can be applied to both halves separately; the residual slot adds the
        # first half's (already scaled and gated) result
        lo = ops.conv_dgrad(dv[..., :nlo], W[f"ups.{i}"][0], (0, -1), gate=a_in, gate_slope=slope, alpha=alpha)
        return ops.conv_dgrad(dv[..., nlo:], W[f"ups.{i}"][1], (1, 0), gate=a_in, gate_slope=slope, alpha=alpha,
                              residual=lo, round_out=True)


class HifiganGenerator(Generator):
    """Drop-in for python/xvapitch/hifigan.py:159 ``HifiganGenerator`` -- the xVAPitch waveform decoder
    (xvapitch/model.py:134-149: 192 latent channels in, resblock '1', upsample factors (8, 8, 2, 2), plain conv_pre /
    conv_post, no conv_post bias, cond_layer d_vector_dim -> 512). Same constructor arguments, parameter names
    (conv_pre.weight, cond_layer.weight, ups.N.weight_g ...) and forward(x, g); the kernels are the HiFi-GAN v1
    generator's. backward(dy, need_input_grad=True) returns the latent's and the conditioning vector's gradients."""

    def __init__(self, in_channels, out_channels, resblock_type, resblock_dilation_sizes, resblock_kernel_sizes,
                 upsample_kernel_sizes, upsample_initial_channel, upsample_factors, inference_padding=5, cond_channels=0,
                 conv_pre_weight_norm=True, conv_post_weight_norm=True, conv_post_bias=True, device=None, seed=1234):
        if out_channels != 1:
            raise NotImplementedError("out_channels != 1: the reference only builds a mono waveform decoder")
        from types import SimpleNamespace
        h = SimpleNamespace(resblock=str(resblock_type), resblock_dilation_sizes=resblock_dilation_sizes,
                            resblock_kernel_sizes=resblock_kernel_sizes, upsample_kernel_sizes=upsample_kernel_sizes,
                            upsample_initial_channel=upsample_initial_channel, upsample_rates=upsample_factors)
        super().__init__(h, device=device, seed=seed, in_channels=in_channels, cond_channels=cond_channels,
                         conv_pre_weight_norm=conv_pre_weight_norm, conv_post_weight_norm=conv_post_weight_norm,
                         conv_post_bias=conv_post_bias)
        self.inference_padding = inference_padding

    def forward(self, x, g=None):
        return super().forward(x, g=g)

    @torch.no_grad()
    def inference(self, c):
        """hifigan.py:264-279: replicate-pad the input by inference_padding frames on both sides, then forward."""
        c = torch.nn.functional.pad(c.to(_wv(self.conv_pre).device), (self.inference_padding,) * 2, "replicate")
        return self.forward(c)


# ------------------------------------------------------------------------------------------------ mel spectrogram
def _slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) of librosa 0.8.1 (Slaney scale, Slaney norm), the call the
    reference makes at meldataset.py:225 -- restated (librosa is a third-party dependency of the reference)."""
    import numpy as np

    if fmax is None:
        fmax = sr / 2.0
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    hz2mel = lambda f: np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)
    mel2hz = lambda m: np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel2hz(np.linspace(hz2mel(np.float64(fmin)), hz2mel(np.float64(fmax)), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return torch.from_numpy(w.astype(np.float32))


class MelSpectrogram:
    """mel_spectrogram(), hifigan/meldataset.py:217-240, with its backward. ``forward(y [B, N])`` returns the log-mel
    in channels-last layout [B, N / hop, n_mels] (transpose(1, 2) of the reference's [B, n_mels, frames]);
    ``backward(dmel)`` returns dL/dy [B, N]. Unlike the reference (which rebuilds the filterbank with librosa on the
    CPU at every call, meldataset.py:224-226) the DFT and mel bases are built once and stay on the device."""

    EPS, LOG_MIN = 1e-9, 1e-5

    def __init__(self, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000,
                 device="cuda", pad=None, mag_eps=None):
        """pad / mag_eps select the variant: the defaults are hifigan/meldataset.py's
        (reflect (n_fft - hop) / 2, sqrt(re^2 + im^2 + 1e-9)); ``MelSpectrogram.tacotron()`` gives the FastPitch dataset's
        TacotronSTFT (reflect n_fft / 2, no epsilon, N / hop + 1 frames)."""
        if n_fft % hop_size or win_size != n_fft:
            raise NotImplementedError("built for win_size == n_fft and n_fft a multiple of hop_size (config_v1.json)")
        self.n_fft, self.hop, self.n_mels = n_fft, hop_size, num_mels
        self.taps = n_fft // hop_size
        self.nb = n_fft // 2 + 1
        self.pad = (n_fft - hop_size) // 2 if pad is None else int(pad)
        if mag_eps is not None:
            self.EPS = float(mag_eps)
        if (2 * self.pad) % hop_size:
            raise NotImplementedError("2 * pad must be a multiple of the hop size")
        self.ld_s = (2 * self.nb + 3) // 4 * 4          # spectrum row stride (16-byte rows)
        self.ld_m = (self.nb + 31) // 32 * 32           # magnitude row stride / padded K of the mel projection
        dev = torch.device(device)
        n = torch.arange(n_fft, dtype=torch.float64)
        window = torch.hann_window(win_size, periodic=True, dtype=torch.float64)
        ang = 2.0 * math.pi * torch.outer(torch.arange(self.nb, dtype=torch.float64), n) / n_fft
        basis = torch.cat([torch.cos(ang), -torch.sin(ang)], 0) * window[None, :]                # [2 nb, n_fft]
        bw = basis.view(2 * self.nb, self.taps, hop_size).permute(1, 0, 2).contiguous().float()  # [taps, 2 nb, hop]
        self.dft = torch.empty_like(bw, device=dev)
        ops.round_tf32_(bw.to(dev).reshape(-1), self.dft.reshape(-1))
        mel = torch.zeros(1, num_mels, self.ld_m)
        mel[0, :, :self.nb] = _slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
        self.mel = torch.empty_like(mel, device=dev)
        ops.round_tf32_(mel.to(dev).reshape(-1), self.mel.reshape(-1))
        self._ctx = None

    def forward(self, y):
        B, N = y.shape
        if N % self.hop:
            raise ValueError(f"signal length {N} is not a multiple of the hop size {self.hop}")
        F_ = (N + 2 * self.pad - self.n_fft) // self.hop + 1                          # N / hop, or N / hop + 1 (tacotron)
        yp = ops.reflect_pad(y.to(torch.float32).contiguous(), self.pad)              # [B, N + 2 pad]
        view = yp.view(B, F_ + self.taps - 1, self.hop)
        spec = torch.empty(B, F_, self.ld_s, device=y.device, dtype=torch.float32)
        if self.ld_s > 2 * self.nb:
            spec[..., 2 * self.nb:].zero_()
        ops.conv_fwd(view, self.dft, tuple(range(self.taps)), out=spec[..., :2 * self.nb], out_rows=F_)
        mag = ops.spec_mag(spec, self.nb, self.ld_m, self.EPS)
        lin = ops.conv_fwd(mag[..., :self.nb], self.mel[..., :self.nb])
        out = ops.log_clamp(lin, self.LOG_MIN)
        self._ctx = (B, N, F_, spec, lin)
        return out

    __call__ = forward

    @classmethod
    def tacotron(cls, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0, device="cuda"):
        """TacotronSTFT(...).mel_spectrogram of fastpitch1_1/common/layers.py:102-138 (+ common/stft.py:86-114), the mel
        extractor of the FastPitch dataset (data_function.py:226-228): same constructor arguments; ``forward(y [B, N])``
        -> [B, N / hop + 1, n_mels] (channels-last; the reference returns its transpose)."""
        return cls(filter_length, n_mel_channels, sampling_rate, hop_length, win_length, mel_fmin, mel_fmax, device=device,
                   pad=filter_length // 2, mag_eps=0.0)

    @classmethod
    def vits(cls, n_fft=1024, hop_length=256, win_length=1024, sample_rate=22050, mel_fmin=0, mel_fmax=8000, n_mels=80,
             device="cuda"):
        """TorchSTFT(n_fft, hop, win, sample_rate=..., use_mel=True, do_amp_to_db=True) of python/xvapitch/audio.py:40-195
        as VitsGeneratorLoss builds it (xvapitch/losses.py:29-46): centred reflect padding (N / hop + 1 frames) and
        sqrt(clamp(re^2 + im^2, 1e-8)) instead of an added epsilon."""
        return cls(n_fft, n_mels, sample_rate, hop_length, win_length, mel_fmin, mel_fmax, device=device, pad=n_fft // 2,
                   mag_eps=-1e-8)

    def backward(self, dmel):
        B, N, F_, spec, lin = self._ctx
        dlin = ops.log_clamp_bwd(dmel.contiguous(), lin, self.LOG_MIN)
        dmag = ops.conv_dgrad(dlin, self.mel)                                        # [B, F, ld_m]
        dspec = ops.spec_mag_bwd(dmag, spec, self.nb, self.EPS)
        dview = ops.conv_dgrad(dspec[..., :2 * self.nb], self.dft, tuple(range(self.taps)), out_rows=F_ + self.taps - 1)
        return ops.reflect_pad_bwd(dview.view(B, -1), N, self.pad)


# ================================================================================================ discriminators
def _round_up(a, b):
    return (a + b - 1) // b * b


class _DiscConv(nn.Module):
    """Parameters of one discriminator convolution with the reference's names: weight-normed (bias, weight_g, weight_v)
    or spectral-normed (bias, weight_orig, weight_u, weight_v; torch.nn.utils.spectral_norm, one power iteration per
    training forward). ``conv2d`` keeps the trailing singleton kernel dimension of DiscriminatorP's Conv2d weights."""

    def __init__(self, cin, cout, k, stride, pad, groups=1, spectral=False, conv2d=False):
        super().__init__()
        self.cin, self.cout, self.k, self.stride, self.pad, self.groups = cin, cout, k, stride, pad, groups
        self.spectral, self.conv2d = spectral, conv2d
        # channels of the activation buffer this layer writes: the real ones, or (first layer of the VITS scale
        # discriminator, 16 channels) padded with zero channels to the 32-column granule of the next layer's operand
        self.cout_phys = cout
        shape = (cout, cin // groups, k) + ((1,) if conv2d else ())
        self.bias = nn.Parameter(torch.zeros(cout))
        if spectral:
            self.weight_orig = nn.Parameter(torch.zeros(shape))
            self.register_buffer("weight_u", torch.zeros(cout))
            self.register_buffer("weight_v", torch.zeros(int(math.prod(shape[1:]))))
        else:
            self.weight_g = nn.Parameter(torch.ones((cout,) + (1,) * (len(shape) - 1)))
            self.weight_v = nn.Parameter(torch.zeros(shape))

    def weight(self):
        """Effective [Cout, Cin/groups, k] weight under autograd (and, in training mode, the spectral-norm power
        iteration on the u / v buffers exactly as torch's forward pre-hook does it)."""
        if not self.spectral:
            v = self.weight_v
            w = v * (self.weight_g / v.flatten(1).norm(dim=1).view((-1,) + (1,) * (v.dim() - 1)))
        else:
            wo = self.weight_orig
            mat = wo.reshape(self.cout, -1)
            u, v = self.weight_u, self.weight_v
            if self.training:
                with torch.no_grad():
                    v = torch.nn.functional.normalize(torch.mv(mat.t(), u), dim=0, eps=1e-12, out=v)
                    u = torch.nn.functional.normalize(torch.mv(mat, v), dim=0, eps=1e-12, out=u)
                    u, v = u.clone(), v.clone()
            sigma = torch.dot(u, torch.mv(mat, v))
            w = wo / sigma
        return w.reshape(self.cout, self.cin // self.groups, self.k)

    # ---- tap geometry of the (strided) convolution on the [L/stride, stride*Cin] view of its input
    def taps(self):
        """[(j, row shift, phase)] ordered by phase so that the taps of one phase are consecutive in the packed weight."""
        t = []
        for j in range(self.k):
            o = j - self.pad
            t.append((j, o // self.stride, o % self.stride))
        return sorted(t, key=lambda e: (e[2], e[0]))

    def out_len(self, L):
        return (L + 2 * self.pad - self.k) // self.stride + 1


class _Disc(nn.Module):
    """One sub-discriminator: convs[0] reads the raw waveform (one channel), convs[1:] and conv_post are tap-GEMMs."""

    def __init__(self, specs, post, period=None, spectral=False):
        super().__init__()
        c2d = period is not None
        self.period = period
        self.convs = nn.ModuleList([_DiscConv(*s, spectral=spectral, conv2d=c2d) for s in specs])
        self.conv_post = _DiscConv(*post, spectral=spectral, conv2d=c2d)
        m0 = self.convs[0]
        if m0.cout % 32:
            if spectral:
                raise NotImplementedError("a spectral-normed first layer with a channel count that is not a multiple of 32")
            m0.cout_phys = _round_up(m0.cout, 32)

    # geometry of the Z = B*P sequences inside the waveform [B, T]
    def _geom(self, T):
        if self.period is None:
            return (T, 1, 0, 1, T, T)
        p = self.period
        return (T, p, 1, p, T, (T + p - 1) // p)

    def _group_geom(self, m):
        """(Gp, Ogp, Cgp): launch-level grouping of a layer. Groups with fewer than 32 input channels are merged f at
        a time into super-groups with a block-diagonal filter (zeros off the diagonal), so that every k-block the
        tensor core reads is 32 real channels wide and one launch covers the whole layer (xva_gemm_args.groups)."""
        G, Og, Cg = m.groups, m.cout // m.groups, m.cin // m.groups
        f = max(1, 32 // Cg) if G > 1 else 1
        if G < f and G * Cg < 32 and 32 % Cg == 0:
            # fewer groups than one 32-column block holds (VITS scale discriminator, 16 -> 64 in 4 groups): ONE dense
            # launch over the input zero-padded to 32 channels (cout_phys of the layer before), block-diagonal filter
            return 1, m.cout, Cg * f, f
        if G % f:
            raise NotImplementedError(f"groups={G} with {Cg} channels per group")
        return G // f, Og * f, Cg * f, f

    def _packed(self):
        """per layer: packed weight [k (phase-major), Cout, Cgp] (autograd) for layers >= 1 and conv_post; layer 0 uses
        its [Cout, k] weight directly."""
        out = []
        for li, m in enumerate(list(self.convs) + [self.conv_post]):
            w = m.weight()
            if li == 0:
                w0 = w.reshape(m.cout, m.k)
                if m.cout_phys != m.cout:
                    w0 = torch.nn.functional.pad(w0, (0, 0, 0, m.cout_phys - m.cout))
                out.append(w0.contiguous())
                continue
            if getattr(m, "_order_idx", None) is None or m._order_idx.device != w.device:
                m._order_idx = torch.tensor([j for j, _, _ in m.taps()], device=w.device, dtype=torch.long)
                Gp, Ogp, Cgp, f = self._group_geom(m)
                Og = m.cout // m.groups
                slot = (torch.arange(m.cout, device=w.device) // Og) % f
                m._slot_mask = (slot[:, None] == torch.arange(f, device=w.device)[None, :]).to(torch.float32)[:, :, None]
            order = m._order_idx          # cached on the device: no host->device copy inside a captured step
            Gp, Ogp, Cgp, f = self._group_geom(m)
            wk = w.index_select(2, order).permute(2, 0, 1)                       # [k, Cout, Cg]
            if f > 1:
                wk = (wk.unsqueeze(2) * m._slot_mask).reshape(m.k, m.cout, Cgp)  # block-diagonal super-groups
            out.append(wk.contiguous())
        return out

    def _lens(self, Z, L, dev):
        key = (Z, L)
        cache = self.__dict__.setdefault("_lens_cache", {})
        if key not in cache:
            cache[key] = torch.full((Z,), L, device=dev, dtype=torch.int32)
        return cache[key]

    def register_weights(self, packer, prefix):
        """Describe this sub-discriminator's convolutions to a _WnPacker (same layouts as _packed())."""
        for li, m in enumerate(list(self.convs) + [self.conv_post]):
            key = f"{prefix}.{li}"
            if li == 0:
                packer.add_flat(key, m, m.cout_phys, m.k)      # rows >= m.cout stay zero (the arena is zero-filled)
            else:
                Gp, Ogp, Cgp, f = self._group_geom(m)
                packer.add_conv(key, m, m.cout, m.cin // m.groups, m.k, order=[j for j, _, _ in m.taps()],
                                og=m.cout // m.groups, f=f)

    def forward(self, wave, keep=True, weight_grad=True, W=None, gW=None):
        """wave [B, T] fp32 -> (score [Z, L, 1], fmaps [activation buffers + score], ctx). W / gW: this
        sub-discriminator's packed weights and their gradient buffers when the model packs all weights in one launch
        (_WnPacker); otherwise they are produced here with PyTorch ops under autograd (the spectral-norm path)."""
        B, T = wave.shape
        geom = self._geom(T)
        P, L = geom[3], geom[5]
        Z = B * P
        if W is None:
            with torch.set_grad_enabled(weight_grad and torch.is_grad_enabled()):
                packed = self._packed()
        else:
            packed = None
        layers = list(self.convs)
        m0 = layers[0]
        L0 = m0.out_len(L)
        nxt = layers[1].stride
        w0 = packed[0].detach() if W is None else W[0]
        b0 = m0.bias.detach()
        if m0.cout_phys != m0.cout:
            pad = self.__dict__.get("_bias_pad")
            if pad is None or pad.device != b0.device:
                pad = self.__dict__["_bias_pad"] = torch.zeros(m0.cout_phys, device=b0.device, dtype=torch.float32)
            pad[:m0.cout].copy_(b0)
            b0 = pad
        X = ops.conv_c1_fwd(wave, geom, w0, b0, m0.k, m0.stride, m0.pad, Z, L0,
                            _round_up(L0, nxt), m0.cout_phys, LRELU_SLOPE)
        acts, lens_v = [X], [L0]
        Wr = [w0]
        for li in range(1, len(layers)):
            m = layers[li]
            Lin = lens_v[-1]
            Lout = m.out_len(Lin)
            s_next = layers[li + 1].stride if li + 1 < len(layers) else 1
            Lp_out = _round_up(Lout, s_next)
            out = torch.empty(Z, Lp_out, m.cout, device=wave.device, dtype=torch.float32)
            wr = _rounded(packed[li]) if W is None else W[li]
            Wr.append(wr)
            self._layer_fwd(m, X, wr, out, Lp_out, self._lens(Z, Lout, wave.device))
            X = out
            acts.append(X)
            lens_v.append(Lout)
        mp = self.conv_post
        wrp = _rounded(packed[-1]) if W is None else W[-1]
        Wr.append(wrp)
        Lf = lens_v[-1]
        score = torch.empty(Z, Lf, 1, device=wave.device, dtype=torch.float32)
        taps = mp.taps()
        ops.conv_fwd(X, wrp, [sh for _, sh, _ in taps], out=score, out_rows=Lf, bias=mp.bias.detach())
        ctx = dict(wave=wave, geom=geom, Z=Z, acts=acts, lens=lens_v, packed=packed, Wr=Wr, score=score, gW=gW) if keep else None
        return score, acts + [score], ctx

    def _layer_fwd(self, m, X, wr, out, Lp_out, lens_t):
        Z, Lp_in, Cin = X.shape
        s = m.stride
        xv = X.view(Z, Lp_in // s, s * Cin)
        taps = m.taps()
        Gp, Ogp, Cgp, _ = self._group_geom(m)
        ops.conv_fwd(xv, wr, [sh for _, sh, _ in taps], a_cols=[ph * Cin for _, _, ph in taps], out=out, out_rows=Lp_out,
                     lens=lens_t, bias=m.bias.detach(), act_slope=LRELU_SLOPE, round_out=True, groups=Gp, grp_step=Cgp)

    @staticmethod
    def slice_ctx(ctx, lo, hi):
        """The part of a batched pass that belongs to sequences [lo, hi) (e.g. the generated half of a real + generated
        batch): views, nothing is copied."""
        P = ctx["geom"][3]
        out = dict(ctx)
        out.update(Z=hi - lo, acts=[a[lo:hi] for a in ctx["acts"]], score=ctx["score"][lo:hi],
                   wave=ctx["wave"][lo // P:hi // P])
        return out

    def backward(self, ctx, dscore, dfeat, need_w, dwave=None, wave_scale=1.0):
        """dscore [Z, L, 1]: gradient wrt the score; dfeat[l]: gradient wrt the PRE-activation of acts[l] coming from the
        feature loss (already gated), or None. Accumulates parameter gradients when need_w, and dL/d(waveform) into dwave
        when given."""
        Z, acts, lens_v, Wr = ctx["Z"], ctx["acts"], ctx["lens"], ctx["Wr"]
        layers = list(self.convs)
        dev = dscore.device
        gW = None
        shared = ctx.get("gW") is not None       # gradient views of the model-level _WnPacker (already zeroed)
        if need_w and shared:
            gW = list(ctx["gW"])
        elif need_w:   # one zero-filled buffer for every packed-weight gradient of this sub-discriminator
            sizes = [w.numel() for w in Wr[1:]]
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
            gW, off = [None], 0
            for w, n in zip(Wr[1:], sizes):
                gW.append(flat[off:off + n].view(w.shape))
                off += n

        def bias_grad(m, d, cols, ld):
            if m.bias.grad is None:
                m.bias.grad = torch.zeros_like(m.bias)
            _Side.run(lambda: ops.colsum_(d.shape[0] * d.shape[1], cols, ld, d, m.bias.grad), d)

        # conv_post: pad the single score channel to 32 columns (MN-major operand rule)
        mp = self.conv_post
        Lf = lens_v[-1]
        dp = torch.zeros(Z, Lf, 32, device=dev, dtype=torch.float32)
        dp[..., 0] = dscore[..., 0]
        ops.round_tf32_(dp.reshape(-1), dp.reshape(-1))
        d1 = dp[..., :1]
        X = acts[-1]
        taps = mp.taps()
        shifts = [sh for _, sh, _ in taps]
        if need_w:
            _Side.run(lambda: ops.conv_wgrad(d1, X, shifts, out=gW[-1], accumulate=True, dy_rows=Lf), d1, X)
            bias_grad(mp, dscore.contiguous(), 1, 1)
        dpre = ops.conv_dgrad(d1, Wr[-1], shifts, out_rows=X.shape[1], gate=X, gate_slope=LRELU_SLOPE,
                              residual=dfeat[len(layers) - 1], lens=self._lens(Z, Lf, dev), round_out=True)
        for li in range(len(layers) - 1, 0, -1):
            m = layers[li]
            Xin = acts[li - 1]
            Zz, Lp_in, Cin = Xin.shape
            s = m.stride
            xv = Xin.view(Z, Lp_in // s, s * Cin)
            taps = m.taps()
            Gp, Ogp, Cgp, _ = self._group_geom(m)
            if need_w:
                bias_grad(m, dpre, m.cout, m.cout)
                _Side.run(lambda dpre=dpre, xv=xv, taps=taps, Cin=Cin, Cgp=Cgp, Gp=Gp, li=li: ops.conv_wgrad(
                    dpre, xv, [sh for _, sh, _ in taps], x_cols=[ph * Cin for _, _, ph in taps], n_cols=Cgp,
                    out=gW[li], accumulate=True, groups=Gp, grp_step=Cgp), dpre, xv)
            if li == 1 and dwave is None and not need_w:
                break
            dX = torch.empty_like(Xin)
            dxv = dX.view(Z, Lp_in // s, s * Cin)
            res = dfeat[li - 1]
            resv = res.view(Z, Lp_in // s, s * Cin) if res is not None else None
            for ph in range(s):
                idx = [i for i, (_, _, p_) in enumerate(taps) if p_ == ph]
                lo, hi = idx[0], idx[-1] + 1
                c0, c1 = ph * Cin, (ph + 1) * Cin
                ops.conv_dgrad(dpre, Wr[li][lo:hi], [taps[i][1] for i in idx], out=dxv[..., c0:c1], out_rows=Lp_in // s,
                               gate=xv[..., c0:c1], gate_slope=LRELU_SLOPE,
                               residual=None if resv is None else resv[..., c0:c1], round_out=True, groups=Gp)
            if Lp_in > lens_v[li - 1]:
                ops.zero_tail_rows_(dX, lens_v[li - 1])
            dpre = dX
        # first layer (raw waveform)
        m0 = layers[0]
        w0 = Wr[0] if shared else ctx["packed"][0]
        if need_w:
            dw0 = gW[0] if shared else torch.zeros(m0.cout_phys, m0.k, device=dev, dtype=torch.float32)
            if m0.bias.grad is None:
                m0.bias.grad = torch.zeros_like(m0.bias)
            db0 = m0.bias.grad if m0.cout_phys == m0.cout else torch.zeros(m0.cout_phys, device=dev, dtype=torch.float32)
            ops.conv_c1_bwd_w(dpre, ctx["wave"], ctx["geom"], m0.k, m0.stride, m0.pad, lens_v[0], dw0, db0)
            if db0 is not m0.bias.grad:
                m0.bias.grad.add_(db0[:m0.cout])
        if dwave is not None:
            ops.conv_c1_bwd_x(dpre, w0.detach(), ctx["geom"], m0.k, m0.stride, m0.pad, lens_v[0], wave_scale, dwave)
        if need_w and not shared:
            _Side.join()
            torch.autograd.backward([w0] + list(ctx["packed"][1:]), [dw0] + gW[1:])


def _rounded(t):
    out = torch.empty_like(t)
    ops.round_tf32_(t.detach().reshape(-1), out.reshape(-1))
    return out


class DiscriminatorP(_Disc):
    """hifigan/models.py:140-173"""

    def __init__(self, period, kernel_size=5, stride=3, use_spectral_norm=False):
        specs = [(1, 32, kernel_size, stride, 2), (32, 128, kernel_size, stride, 2), (128, 512, kernel_size, stride, 2),
                 (512, 1024, kernel_size, stride, 2), (1024, 1024, kernel_size, 1, 2)]
        super().__init__(specs, (1024, 1, 3, 1, 1), period=period, spectral=use_spectral_norm)


class DiscriminatorS(_Disc):
    """hifigan/models.py:203-228"""

    def __init__(self, use_spectral_norm=False):
        specs = [(1, 128, 15, 1, 7), (128, 128, 41, 2, 20, 4), (128, 256, 41, 2, 20, 16), (256, 512, 41, 4, 20, 16),
                 (512, 1024, 41, 4, 20, 16), (1024, 1024, 41, 1, 20, 16), (1024, 1024, 5, 1, 2)]
        super().__init__(specs, (1024, 1, 3, 1, 1), period=None, spectral=use_spectral_norm)


class MultiPeriodDiscriminator(nn.Module):
    """hifigan/models.py:176-200. forward(y, y_hat) -> (y_d_rs, y_d_gs, fmap_rs, fmap_gs) with channels-last fmaps
    [B*p, L, C]; ctxs are kept on the module for backward."""

    def __init__(self, device=None, seed=1234):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorP(p) for p in (2, 3, 5, 7, 11)])
        _init_disc(self, seed, device)

    def forward(self, y, y_hat, weight_grad=True, join=True, stream_offset=0):
        return _multi_forward(self, y, y_hat, pools=0, weight_grad=weight_grad, join=join, stream_offset=stream_offset)


class MultiScaleDiscriminator(nn.Module):
    """hifigan/models.py:231-260"""

    def __init__(self, device=None, seed=1234):
        super().__init__()
        sn = os.environ.get("XVA_DEBUG_NO_SPECTRAL", "0") != "1"     # diagnostic only: bounds what the spectral-norm path costs
        self.discriminators = nn.ModuleList([DiscriminatorS(use_spectral_norm=sn), DiscriminatorS(), DiscriminatorS()])
        _init_disc(self, seed + 1, device)

    def forward(self, y, y_hat, weight_grad=True, join=True, stream_offset=0):
        return _multi_forward(self, y, y_hat, pools=1, weight_grad=weight_grad, join=join, stream_offset=stream_offset)


class VitsDiscriminatorS(_Disc):
    """python/xvapitch/model.py:1548-1587: the scale discriminator with VITS channel widths (16, 64, 256, 1024, 1024,
    1024; four input channels per group). The 16-channel first layer writes a 32-channel buffer with zero upper half,
    and groups are merged eight at a time into block-diagonal 32-column super-groups (_group_geom)."""

    def __init__(self, use_spectral_norm=False):
        specs = [(1, 16, 15, 1, 7), (16, 64, 41, 4, 20, 4), (64, 256, 41, 4, 20, 16), (256, 1024, 41, 4, 20, 64),
                 (1024, 1024, 41, 4, 20, 256), (1024, 1024, 5, 1, 2)]
        super().__init__(specs, (1024, 1, 3, 1, 1), period=None, spectral=use_spectral_norm)


class VitsDiscriminator(nn.Module):
    """Drop-in for python/xvapitch/model.py:1590 ``VitsDiscriminator``: nets[0] the scale discriminator above, nets[1:]
    DiscriminatorP(2, 3, 5, 7, 11) (xvapitch/hifigan.py:301-370 -- the layer table of hifigan/models.py's). Same
    state_dict keys (nets.N.convs.M.weight_g ...). forward(x, x_hat) -> (x_scores, x_feats, x_hat_scores, x_hat_feats)
    in the reference's order, feature maps channels-last [B*p, L, C]; discriminator_loss_backward /
    generator_adv_loss_backward (pools=0) are its losses (xvapitch/losses.py:65-85, 329-342 restate
    hifigan/models.py:263-294)."""

    def __init__(self, use_spectral_norm=False, device=None, seed=1234):
        super().__init__()
        if use_spectral_norm:
            raise NotImplementedError("the reference builds VitsDiscriminator(use_spectral_norm=False) (xvapitch/model.py:151)")
        self.nets = nn.ModuleList([VitsDiscriminatorS()] + [DiscriminatorP(p) for p in (2, 3, 5, 7, 11)])
        _init_disc(self, seed + 2, device)

    @property
    def discriminators(self):
        return self.nets

    def forward(self, x, x_hat=None, weight_grad=True):
        if x_hat is None:
            rs, _, frs, _ = _multi_forward(self, x, x, pools=0, weight_grad=weight_grad)
            return rs, frs, None, None
        rs, gs, frs, fgs = _multi_forward(self, x, x_hat, pools=0, weight_grad=weight_grad)
        return rs, frs, gs, fgs


def _init_disc(model, seed, device):
    g = torch.Generator().manual_seed(int(seed))
    with torch.no_grad():
        for m in model.modules():
            if not isinstance(m, _DiscConv):
                continue
            w = m.weight_orig if m.spectral else m.weight_v
            fan_in = int(math.prod(w.shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * bound)
            m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
            if m.spectral:
                m.weight_u.copy_(torch.nn.functional.normalize(torch.randn(m.weight_u.shape, generator=g), dim=0))
                m.weight_v.copy_(torch.nn.functional.normalize(torch.randn(m.weight_v.shape, generator=g), dim=0))
            else:
                m.weight_g.copy_(w.flatten(1).norm(dim=1).view(m.weight_g.shape))
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if dev.type != "cuda":
        raise capi.XvaError("the discriminators (B200 build) need a CUDA device: there is no CPU path")
    capi.load()
    model.to(dev)


def _multi_forward(model, y, y_hat, pools, weight_grad=True, join=True, stream_offset=0):
    """join=False leaves the branch streams open (the caller joins after it has issued another model's branches too:
    MPD and MSD are independent, HiFiGANStep runs their eight sub-discriminators side by side on streams
    stream_offset + i); the waveforms the branches read then stay referenced on the model until its next forward.
    Shared by MPD / MSD: every sub-discriminator on the real and the generated waveform ([B, 1, T] or [B, T]).
    Weight-normed sub-discriminators see both waveforms as ONE batch of 2B sequences (one launch per layer instead of
    two); the spectral-normed one (MSD scale 0) runs them one after the other because torch's spectral_norm hook does a
    power iteration per call, so the reference's two calls (models.py:251-252) use two different weights.
    model._ctx[i] = list of passes (ctx, rows of the real sequences or None, rows of the generated ones or None)."""
    yr = y.reshape(y.shape[0], -1).to(torch.float32)
    yg = y_hat.reshape(y_hat.shape[0], -1).to(torch.float32)
    B = yr.shape[0]
    both = torch.cat([yr, yg], 0)
    model._ctx = []
    pk = getattr(model, "_packer", None)
    if pk is None:
        pk = _WnPacker()
        for i, d in enumerate(model.discriminators):
            if not any(m.spectral for m in d.convs):
                d.register_weights(pk, str(i))
        pk.finalize(both.device)
        model._packer = pk
    Wall = pk.pack()                      # every weight-normed convolution of the model: one launch
    sn = getattr(model, "_sn", None)
    if sn is None and os.environ.get("XVA_SN_NATIVE", "1") != "0" and any(m.spectral for d in model.discriminators for m in d.convs):
        sn = _SnPacker(slots=2)
        for i, d in enumerate(model.discriminators):
            if any(m.spectral for m in d.convs):
                d.register_weights(sn, str(i))
        sn.finalize(both.device)
        model._sn = sn
    n_layers = lambda d: len(d.convs) + 1
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    inputs = [both]                       # every pooled waveform stays referenced until the branches have been joined
    for i, d in enumerate(model.discriminators):
        if pools and i != 0:
            both = ops.avgpool4(both)
            inputs.append(both)
        with _Branches.branch(i + stream_offset):
            if any(m.spectral for m in d.convs) and sn is not None:
                # one power iteration + pack per call, real first (models.py:251-252), each call with its own weights
                outs = []
                for slot, part in ((0, both[:B]), (1, both[B:])):
                    Ws = sn.pack(slot, model.training)
                    outs.append(d(part, weight_grad=weight_grad, W=[Ws[f"{i}.{li}"][0] for li in range(n_layers(d))],
                                  gW=[sn.gW[slot][f"{i}.{li}"][0] for li in range(n_layers(d))]))
                (sr, fr, cr), (sg, fg, cg) = outs
                passes = [(cr, (0, cr["Z"]), None), (cg, None, (0, cg["Z"]))]
            elif any(m.spectral for m in d.convs):
                sr, fr, cr = d(both[:B], weight_grad=weight_grad)
                sg, fg, cg = d(both[B:], weight_grad=weight_grad)
                passes = [(cr, (0, cr["Z"]), None), (cg, None, (0, cg["Z"]))]
            else:
                s2, f2, c2 = d(both, weight_grad=weight_grad, W=[Wall[f"{i}.{li}"][0] for li in range(n_layers(d))],
                               gW=[pk.gW[f"{i}.{li}"][0] for li in range(n_layers(d))])
                Z = c2["Z"] // 2
                sr, sg = s2[:Z], s2[Z:]
                fr, fg = [f[:Z] for f in f2], [f[Z:] for f in f2]
                passes = [(c2, (0, Z), (Z, 2 * Z))]
        y_d_rs.append(sr)
        y_d_gs.append(sg)
        fmap_rs.append(fr)
        fmap_gs.append(fg)
        model._ctx.append(passes)
    model._branch_inputs = inputs
    if join:
        _Branches.join()
        model._branch_inputs = None
    del inputs
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


# ================================================================================================ losses / steps
def discriminator_loss_backward(model, y_d_rs, y_d_gs, join=True, stream_offset=0):
    """discriminator_loss (models.py:272-283) on the outputs of ``model(y, y_hat.detach())`` plus the backward into the
    discriminator's parameters (hifigan/xva_train.py:486-497). Returns the loss as a 0-dim device tensor -- or, with
    join=False, a function that returns it and must be called after _Branches.join() (it sums the per-branch terms and
    unpacks the weight gradients, both of which need every branch finished)."""
    dev = y_d_rs[0].device
    acc = torch.zeros(2 * len(y_d_rs), device=dev, dtype=torch.float64)
    model._packer.zero_grads()
    if getattr(model, "_sn", None) is not None:
        model._sn.zero_grads()
    parts = []
    for i, (d, passes) in enumerate(zip(model.discriminators, model._ctx)):
        with _Branches.branch(i + stream_offset):
            dr, dg = y_d_rs[i], y_d_gs[i]
            n = dr.numel()
            ops.reduce_sq(dr, 1.0, acc[2 * i:2 * i + 1])
            ops.reduce_sq(dg, 0.0, acc[2 * i + 1:2 * i + 2])
            parts.append((acc[2 * i] + acc[2 * i + 1]) / n)
            for ctx, rr, gr in passes:
                dscore = torch.empty_like(ctx["score"])
                if rr is not None:
                    ops.sq_grad(dr, 1.0, 1.0 / n, out=dscore[rr[0]:rr[1]], accumulate=False)
                if gr is not None:
                    ops.sq_grad(dg, 0.0, 1.0 / n, out=dscore[gr[0]:gr[1]], accumulate=False)
                d.backward(ctx, dscore, [None] * len(ctx["acts"]), need_w=True)

    def finish():
        loss = torch.zeros((), device=dev, dtype=torch.float64)
        for part in parts:                # same summation order as one sub-discriminator after the other
            loss = loss + part
        model._packer.unpack_grads()      # packed-weight gradients -> weight_g / weight_v of every sub-discriminator
        if getattr(model, "_sn", None) is not None:
            model._sn.unpack_grads()      # ... and -> weight_orig of the spectral-normed one
        model._branch_inputs = None
        return loss

    if not join:
        return finish
    _Branches.join()
    return finish()


def generator_adv_loss_backward(model, y_d_gs, fmap_rs, fmap_gs, dwave, pools, fm_grad=True, join=True, stream_offset=0):
    """generator_loss + feature_loss (models.py:263-269, 286-294) on the outputs of ``model(y, y_hat)`` and their
    gradient wrt the generated waveform, ACCUMULATED into dwave [B, T] (hifigan/xva_train.py:506-513). The
    discriminator weights get no gradient here: the reference computes and then discards it (its zero_grad at
    :468-469 / :483 clears it before any optimizer step reads it). fm_grad=False: the feature-matching term is only
    evaluated -- xVAPitch's generator loss detaches the generated features (python/xvapitch/losses.py:196 passes
    (fake, real) to feature_loss(feats_real, feats_generated), which detaches its first argument, :69). join=False: returns
    a function to call after _Branches.join() that finishes the accumulation into dwave and returns the two losses."""
    dev = dwave.device
    n_d = len(y_d_gs)
    acc = torch.zeros(n_d * 16, device=dev, dtype=torch.float64)
    levels = [dwave]
    for i in range(1, n_d):
        if pools:
            L = levels[-1].shape[1]
            levels.append(torch.zeros(dwave.shape[0], L // 2 + 1, device=dev, dtype=torch.float32))
    gen_parts, fm_parts, own_dwave = [], [], []
    for i in reversed(range(n_d)):
        with _Branches.branch(i + stream_offset):
            d = model.discriminators[i]
            cg = None
            for ctx, rr, gr in model._ctx[i]:      # the pass (or the half of the batched pass) of the generated waveform
                if gr is not None:
                    cg = ctx if rr is None else _Disc.slice_ctx(ctx, gr[0], gr[1])
            dg = y_d_gs[i]
            n = dg.numel()
            ops.reduce_sq(dg, 1.0, acc[16 * i:16 * i + 1])
            gen_parts.append(acc[16 * i] / n)
            dscore = ops.sq_grad(dg, 1.0, 1.0 / n)
            dfeat = []
            n_act = len(cg["acts"])
            for l, (fr, fg) in enumerate(zip(fmap_rs[i], fmap_gs[i])):
                # elements of the reference's feature map (zero-padded channels of the buffer are not part of it)
                valid = cg["Z"] * (cg["lens"][l] if l < n_act else cg["lens"][-1]) * (d.convs[l].cout if l < n_act else 1)
                a = acc[16 * i + 1 + l:16 * i + 2 + l]
                if not fm_grad:
                    ops.reduce_l1(fr, fg, a)
                    if l < n_act:
                        dfeat.append(None)
                elif l < n_act:
                    dfeat.append(ops.l1_loss_grad(fr, fg, 2.0 / valid, a, gate_slope=LRELU_SLOPE))   # loss term + gradient
                else:
                    ops.reduce_l1(fr, fg, a)
                    ops.l1_grad(fr, fg, 2.0 / valid, out=dscore)     # conv_post output is the last feature map too
                fm_parts.append(2.0 * a[0] / valid)
            if pools:
                tgt = levels[i]
            elif _Branches.on():                   # concurrent branches must not accumulate into one tensor
                tgt = torch.zeros_like(dwave)
                own_dwave.append(tgt)
            else:
                tgt = dwave
            d.backward(cg, dscore, dfeat, need_w=False, dwave=tgt)

    def finish():
        loss_gen = torch.zeros((), device=dev, dtype=torch.float64)
        loss_fm = torch.zeros((), device=dev, dtype=torch.float64)
        for part in gen_parts:            # same summation order as one sub-discriminator after the other
            loss_gen = loss_gen + part
        for part in fm_parts:
            loss_fm = loss_fm + part
        for t in own_dwave:
            dwave.add_(t)
        if pools:
            for i in reversed(range(1, n_d)):
                up = ops.avgpool4_bwd(levels[i], levels[i - 1].shape[1])
                levels[i - 1].add_(up)
        model._branch_inputs = None
        return loss_gen, loss_fm

    if not join:
        return finish
    _Branches.join()
    return finish()


class AdamW:
    """torch.optim.AdamW(params, lr, betas) of hifigan/xva_train.py:298-300 as ONE launch over a flat arena: at
    construction the parameters are moved into one contiguous fp32 buffer (they become views of it) and their .grad
    fields are bound to views of a second one, so backward writes the arena directly."""

    def __init__(self, params, lr=2e-4, betas=(0.8, 0.99), eps=1e-8, weight_decay=0.01):
        self.params = [p for p in params]
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        dev = self.params[0].device
        # every parameter starts on a 16-byte boundary of the arena (a 1-element bias would otherwise misalign everything
        # behind it and push the weight-pack kernels onto their scalar path); the pad elements stay zero
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)   # device-side step count (graph replay)
        self.lr_on_device = False   # True: lr_dev is maintained by the caller (CUDA-graph replay), step() leaves it
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                k = p.numel()
                self.p[off:off + k].copy_(p.reshape(-1))
                p.data = self.p[off:off + k].view(p.shape)
                p.grad = self.g[off:off + k].view(p.shape)
        self.steps = 0

    def zero_grad(self, set_to_none=False):
        self.g.zero_()
        for p, off in zip(self.params, self.offsets):       # re-bind in case something replaced a .grad tensor
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.g[off:off + k].data_ptr():
                p.grad = self.g[off:off + k].view(p.shape)

    def state_dict(self):
        """torch.optim.AdamW.state_dict() layout: {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}
        with parameter i = the i-th parameter handed to the constructor, tensors on the CPU in the parameters' shapes."""
        st = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            k = p.numel()
            if self.steps > 0:
                st[i] = {"step": torch.tensor(float(self.steps)), "exp_avg": self.m[off:off + k].view(p.shape).cpu().clone(),
                         "exp_avg_sq": self.v[off:off + k].view(p.shape).cpu().clone()}
        g = self.param_groups[0]
        group = {"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": g["weight_decay"],
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(self.params)))}
        return {"state": st, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts torch.optim.AdamW's state_dict (what the reference's do_ checkpoints hold) or this class's round-1 flat
        layout ({'m', 'v', 'steps'})."""
        if "m" in sd and "v" in sd and "state" not in sd:      # round-1 layout: parameters packed without padding
            src = 0
            for p, off in zip(self.params, self.offsets):
                k = p.numel()
                self.m[off:off + k].copy_(sd["m"][src:src + k])
                self.v[off:off + k].copy_(sd["v"][src:src + k])
                src += k
            self.steps = int(sd["steps"])
        else:
            state = sd["state"]
            if len(state) not in (0, len(self.params)):
                raise ValueError(f"optimizer state holds {len(state)} parameters, this model has {len(self.params)}")
            steps = 0
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                k = p.numel()
                e = state.get(i, state.get(str(i)))
                if e is not None:
                    if tuple(e["exp_avg"].shape) != tuple(p.shape):
                        raise ValueError(f"optimizer state {i}: shape {tuple(e['exp_avg'].shape)} != {tuple(p.shape)}")
                    self.m[off:off + k].copy_(e["exp_avg"].reshape(-1).to(torch.float32))
                    self.v[off:off + k].copy_(e["exp_avg_sq"].reshape(-1).to(torch.float32))
                    steps = max(steps, int(float(e["step"])))
            self.steps = steps
            if sd.get("param_groups"):
                g0 = sd["param_groups"][0]
                for key in ("lr", "betas", "eps", "weight_decay"):
                    if key in g0:
                        self.param_groups[0][key] = tuple(g0[key]) if key == "betas" else g0[key]
        self.step_dev.fill_(self.steps)

    def step(self):
        g = self.param_groups[0]
        self.steps += 1
        if not self.lr_on_device:
            self.lr_dev.fill_(float(g["lr"]))
        ops.counter_add_(self.step_dev, 1)
        ops.adamw_step_(self.p, self.g, self.m, self.v, self.lr_dev, g["betas"][0], g["betas"][1], g["eps"],
                        g["weight_decay"], self.steps, self.step_dev)


class HiFiGANStep:
    """The body of HiFiTrainer.iteration (hifigan/xva_train.py:467-515): generator forward, mel of the generated audio,
    D step (MPD + MSD on (y, y_hat.detach()) -> discriminator_loss -> AdamW), G step (45 * L1 mel + feature + adversarial
    losses -> AdamW). ``step(x, y, y_mel)`` takes the reference's batch tensors (x [B, 80, T] input mel, y [B, 8192]
    audio, y_mel [B, 80, T] loss mel) and returns the losses as 0-dim device tensors (no host sync)."""

    def __init__(self, generator, mpd, msd, h, lr=None, betas=None, world=1, group=None):
        """world > 1: one process per GPU, each on its own shard of the utterance batch; the two gradient arenas are
        all-reduced (mean) right before their optimizer steps -- the only collectives on the path."""
        self.generator, self.mpd, self.msd, self.h = generator, mpd, msd, h
        self.world, self.group = int(world), group
        lr = h.learning_rate if lr is None else lr
        betas = (h.adam_b1, h.adam_b2) if betas is None else betas
        dev = next(generator.parameters()).device
        self.mel = MelSpectrogram(h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax_for_loss,
                                  device=dev)
        self.optim_g = AdamW(generator.parameters(), lr, betas)
        self.optim_d = AdamW(list(msd.parameters()) + list(mpd.parameters()), lr, betas)   # itertools.chain order of :300
        self.steps = 0

    def _all_reduce(self, flat_grad):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            flat_grad.mul_(1.0 / self.world)

    def step(self, x, y, y_mel):
        G, mpd, msd = self.generator, self.mpd, self.msd
        B = y.shape[0]
        y = y.reshape(B, -1).to(torch.float32).contiguous()
        with ops.nvtx("hifigan.generator.fwd"):
            y_g_hat = G(x)                                                   # :479
            wave = y_g_hat.reshape(B, -1)
            mel_hat = self.mel(wave)                                         # :480, channels-last [B, F, 80]
        mel_tgt = y_mel.to(torch.float32).transpose(1, 2).contiguous()

        # ---- discriminators (:483-498); y_hat is detached: no gradient reaches the generator here
        self.optim_d.zero_grad()
        # MPD and MSD are independent: their 5 + 3 sub-discriminators are issued side by side (streams 0-4 and 5-7 when
        # the branch streams are on, i.e. inside a captured graph) and joined once
        with ops.nvtx("hifigan.d_step"):
            rs_f, gs_f, _, _ = mpd(y, wave, join=False)
            rs_s, gs_s, _, _ = msd(y, wave, join=False, stream_offset=5)
            fin_f = discriminator_loss_backward(mpd, rs_f, gs_f, join=False)
            fin_s = discriminator_loss_backward(msd, rs_s, gs_s, join=False, stream_offset=5)
            _Branches.join()
            loss_disc_f, loss_disc_s = fin_f(), fin_s()
            self._all_reduce(self.optim_d.g)
            self.optim_d.step()

        # ---- generator (:501-515)
        self.optim_g.zero_grad()
        n = mel_hat.numel()
        acc = torch.zeros(1, device=y.device, dtype=torch.float64)
        ops.reduce_l1(mel_tgt, mel_hat, acc)
        loss_mel = 45.0 * acc[0] / n
        dwave = self.mel.backward(ops.l1_grad(mel_tgt, mel_hat, 45.0 / n))
        with ops.nvtx("hifigan.g_step.discriminators"):
            _, gs_f, frs_f, fgs_f = mpd(y, wave, weight_grad=False, join=False)
            _, gs_s, frs_s, fgs_s = msd(y, wave, weight_grad=False, join=False, stream_offset=5)
            fin_f = generator_adv_loss_backward(mpd, gs_f, frs_f, fgs_f, dwave, pools=False, join=False)
            fin_s = generator_adv_loss_backward(msd, gs_s, frs_s, fgs_s, dwave, pools=True, join=False, stream_offset=5)
            _Branches.join()
            (loss_gen_f, loss_fm_f), (loss_gen_s, loss_fm_s) = fin_f(), fin_s()
        with ops.nvtx("hifigan.g_step.generator_bwd"):
            G.backward(dwave.view(B, 1, -1))
            self._all_reduce(self.optim_g.g)
            self.optim_g.step()
        self.steps += 1
        loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
        return {"loss_gen_all": loss_gen_all, "loss_disc_all": loss_disc_s + loss_disc_f, "loss_mel": loss_mel,
                "mel_error": loss_mel / 45.0, "loss_fm": loss_fm_s + loss_fm_f, "loss_gen": loss_gen_s + loss_gen_f}
