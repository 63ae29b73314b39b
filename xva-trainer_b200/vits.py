"""xVAPitch ``--hifi_only`` training path on the B200 engine: the posterior encoder (a
16-layer WaveNet stack) and the waveform decoder trained against the VITS discriminator.

    WN                 drop-in for python/xvapitch/wavenet.py:16        (same constructor, parameter names, forward)
    PosteriorEncoder   drop-in for python/xvapitch/model.py:1422        (same constructor, parameter names, forward)
    HifiOnlyStep       xVAPitch.train_hifi_only + forward + the trainer's iteration for --hifi_only
                       (model.py:650-678, 271-340, 385-399; xva_train.py:651-736; training_util.py:31-32, 66-67)

The waveform decoder and the discriminator are hifigan.HifiganGenerator / hifigan.VitsDiscriminator. Every convolution is
the tcgen05 tap-GEMM of libxva_b200.so (activations channels-last [B, T, C]); the conditioning slice of each WaveNet
layer rides in the GEMM epilogue's residual slot with row stride 0, the mask of each layer in its ``lens`` slot, the
tf32 operand copy of each fp32 stream in its second output. There is no CPU path: the module raises without the library.
"""
import math

import torch
import torch.nn as nn

from . import capi, ops
from .hifigan import (AdamW, MelSpectrogram, _PlainConv, _Side, _WNConv, _WnPacker, _bias, discriminator_loss_backward,
                      generator_adv_loss_backward)

SPEC_SEGMENT, HOP = 32, 256          # xvapitch/model.py:76, :662-663


def _need_cuda(device):
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if dev.type != "cuda":
        raise capi.XvaError("the xVAPitch modules (B200 build) need a CUDA device: there is no CPU path")
    capi.load()
    capi.call("xva_device_check", dev.index or 0)
    return dev


class WN(nn.Module):
    """python/xvapitch/wavenet.py:16-106. ``forward(x [B, H, T], x_mask, g)`` as there (x_mask [B, 1, T] must be a prefix
    mask: it is turned into lengths); the engine-facing entry points are ``forward_cl`` / ``backward_cl`` on
    channels-last tensors with the packed weights of the owning module's _WnPacker."""

    def __init__(self, in_channels, hidden_channels, kernel_size, dilation_rate, num_layers, c_in_channels=0, dropout_p=0,
                 weight_norm=True):
        super().__init__()
        assert kernel_size % 2 == 1 and hidden_channels % 2 == 0
        if dropout_p != 0 or not weight_norm:
            raise NotImplementedError("dropout_p != 0 / weight_norm=False: not a configuration xVAPitch builds")
        if hidden_channels % 32 or (c_in_channels % 32):
            raise NotImplementedError("hidden / conditioning channels must be multiples of 32")
        self.in_channels, self.hidden_channels, self.kernel_size = in_channels, hidden_channels, kernel_size
        self.dilation_rate, self.num_layers, self.c_in_channels, self.dropout_p = dilation_rate, num_layers, c_in_channels, dropout_p
        H = hidden_channels
        self.in_layers = nn.ModuleList([_WNConv(H, 2 * H, kernel_size, dilation_rate ** i) for i in range(num_layers)])
        self.res_skip_layers = nn.ModuleList([_WNConv(H, 2 * H if i < num_layers - 1 else H, 1) for i in range(num_layers)])
        if c_in_channels > 0:
            self.cond_layer = _WNConv(c_in_channels, 2 * H * num_layers, 1)
        self._ctx = None

    def register_weights(self, packer, prefix):
        for name, m in self.named_modules():
            if isinstance(m, _WNConv):
                packer.add_conv(f"{prefix}{name}", m, m.cout, m.cin, m.k)

    # ------------------------------------------------------------------------------------------ forward
    def forward_cl(self, x, xr, lens, g, W, prefix, keep=True):
        """x [B, T, H] fp32 (already masked), xr its tf32 copy, lens int32 [B], g [1, B, C] tf32 or None ->
        (output * mask [B, T, H], its tf32 copy)."""
        B, T, H = x.shape
        L = self.num_layers
        gcond = None
        if g is not None:
            cl = self.cond_layer
            gcond = ops.conv_fwd(g, W[f"{prefix}cond_layer"][0], (0,), bias=cl.bias.detach())       # [1, B, 2 H L]
        out = out_r = None
        saved = []
        for i in range(L):
            m_in, m_rs = self.in_layers[i], self.res_skip_layers[i]
            res = None
            if gcond is not None:      # g_l broadcast over the frames: residual rows with stride 0 (wavenet.py:97-99)
                res = gcond[0, :, i * 2 * H:(i + 1) * 2 * H].unsqueeze(1).expand(B, T, 2 * H)
            x_in = ops.conv_fwd(xr, W[f"{prefix}in_layers.{i}"][0], m_in.shifts, bias=m_in.bias.detach(), residual=res)
            acts = ops.gated_act(x_in, H)
            w_rs, b_rs = W[f"{prefix}res_skip_layers.{i}"][0], m_rs.bias.detach()
            saved.append((xr, x_in, acts))
            if i < L - 1:
                x_new, xr_new = torch.empty_like(x), torch.empty_like(x)
                ops.conv_fwd(acts, w_rs[:, :H], (0,), out=x_new, bias=b_rs[:H], residual=x, lens=lens, out_act=xr_new,
                             out_act_slope=1.0)                                   # x = (x + res) * mask, :101
                nxt = torch.empty_like(x)
                ops.conv_fwd(acts, w_rs[:, H:], (0,), out=nxt, bias=b_rs[H:], residual=out)          # output += skip, :102
                x, xr, out = x_new, xr_new, nxt
            else:
                nxt, out_r = torch.empty_like(x), torch.empty_like(x)
                ops.conv_fwd(acts, w_rs, (0,), out=nxt, bias=b_rs, residual=out, lens=lens, out_act=out_r,
                             out_act_slope=1.0)                                   # output + res_skip, then * mask, :104-105
                out = nxt
        self._ctx = (saved, lens, g, W, prefix) if keep else None
        return out, out_r

    # ------------------------------------------------------------------------------------------ backward
    def backward_cl(self, d_rs, d_res, gW, bias_grad, wgrad, mask_input_grad=True):
        """d_rs [B, T, 2H]: its upper half holds dL/d(output) (masked, tf32) on entry; d_res [B, T, 2H] is fp32 scratch
        of the same shape. On return the lower half of d_rs (tf32) / d_res (fp32) holds dL/dx of the stack's input
        (times the mask when mask_input_grad: the caller's input was x * mask). Parameter gradients go to gW (packed)
        through ``wgrad`` and to the biases through ``bias_grad``. Returns (dL/dg [1, B, C] (tf32) or None, the tf32 copy
        of dL/dx [B, T, H] -- a view of d_rs or of its alternate buffer)."""
        saved, lens, g, W, prefix = self._ctx
        B, T, H2 = d_rs.shape
        H, L = H2 // 2, self.num_layers
        dG = None
        if g is not None:
            dG = torch.zeros(1, B, 2 * H * L, device=d_rs.device, dtype=torch.float32)
        dxm = d_res[..., :H]
        # With the weight-gradient side stream on (inside a captured graph) the weight / bias gradients of layer i read the
        # gradient buffer of layer i while the launching stream already produces layer i - 1's: two buffers take turns
        # (both carry dL/d(output) in their upper half), and a buffer is rewritten only after the side-stream work that
        # read it two layers earlier has finished (an event; one whole layer of slack).
        side = _Side.on()
        bufs = [d_rs]
        if side:
            alt = torch.empty_like(d_rs)
            alt[..., H:].copy_(d_rs[..., H:])
            bufs.append(alt)
        read_done = [None] * len(bufs)
        cur = 0
        for i in reversed(range(L)):
            xr, x_in, acts = saved[i]
            m_in, m_rs = self.in_layers[i], self.res_skip_layers[i]
            src, nxt = bufs[cur], (cur + 1) % len(bufs)
            d_i = src if i < L - 1 else src[..., H:]            # gradient of res_skip_layers[i]'s output
            bias_grad(m_rs, d_i, m_rs.cout)
            wgrad(d_i, acts, (0,), gW[f"{prefix}res_skip_layers.{i}"][0])
            d_acts = ops.conv_dgrad(d_i, W[f"{prefix}res_skip_layers.{i}"][0], (0,))
            d_xin = ops.gated_act_bwd(d_acts, x_in, H)
            if dG is not None:
                ops.colsum_items_(d_xin, dG[0, :, i * 2 * H:(i + 1) * 2 * H])
            bias_grad(m_in, d_xin, m_in.cout)
            wgrad(d_xin, xr, m_in.shifts, gW[f"{prefix}in_layers.{i}"][0])
            if side:
                read_done[cur] = torch.cuda.Event()
                read_done[cur].record(_Side.stream)
                if read_done[nxt] is not None:
                    torch.cuda.current_stream().wait_event(read_done[nxt])
            # dL/dx_i = dgrad + (the residual path of layers < L-1), then the mask that produced x_i
            ops.conv_dgrad(d_xin, W[f"{prefix}in_layers.{i}"][0], m_in.shifts, out=dxm, residual=dxm if i < L - 1 else None,
                           lens=lens if (i > 0 or mask_input_grad) else None, out_act=bufs[nxt][..., :H], out_act_slope=1.0)
            cur = nxt
        dg = None
        if dG is not None:
            cl = self.cond_layer
            dGr = torch.empty_like(dG)
            ops.round_tf32_(dG.reshape(-1), dGr.reshape(-1))
            bias_grad(cl, dG, cl.cout)
            wgrad(dGr, g, (0,), gW[f"{prefix}cond_layer"][0])
            dg = ops.conv_dgrad(dGr, W[f"{prefix}cond_layer"][0], (0,))
        self._ctx = None
        return dg, bufs[cur][..., :H]

    # ------------------------------------------------------------------------------------------ stand-alone use
    def _own_packer(self):
        if getattr(self, "_packer", None) is None:
            pk = _WnPacker()
            self.register_weights(pk, "")
            pk.finalize(self.in_layers[0].weight_v.device)
            self._packer = pk
        return self._packer

    def forward(self, x, x_mask=None, g=None, **kwargs):
        """WN.forward of the reference (wavenet.py:87-106): x [B, H, T], x_mask [B, 1, T] (a prefix mask) or None,
        g [B, C, 1] or None -> [B, H, T]."""
        B, H, T = x.shape
        dev = self.in_layers[0].weight_v.device
        if x_mask is None:
            lens = torch.full((B,), T, device=dev, dtype=torch.int32)
        else:
            lens = x_mask.reshape(B, T).to(dev).sum(1).to(torch.int32)
        W = self._own_packer().pack()
        xc = x.to(device=dev, dtype=torch.float32).transpose(1, 2).contiguous()
        xr = torch.empty_like(xc)
        ops.round_tf32_(xc.reshape(-1), xr.reshape(-1))
        gr = None
        if g is not None and self.c_in_channels > 0:
            gr = torch.empty(1, B, self.c_in_channels, device=dev, dtype=torch.float32)
            ops.round_tf32_(g.to(device=dev, dtype=torch.float32).reshape(-1).contiguous(), gr.reshape(-1))
        out, _ = self.forward_cl(xc, xr, lens, gr, W, "", keep=self.training)
        return out.transpose(1, 2)

    def backward(self, d_out):
        """d_out = dL/d(output) [B, H, T] -> (dL/dx [B, H, T], dL/dg [B, C, 1] or None); parameter gradients accumulated."""
        if self._ctx is None:
            raise RuntimeError("backward() needs a forward() in training mode first")
        lens = self._ctx[1]
        pk = self._own_packer()
        gW = pk.zero_grads()
        B, H, T = d_out.shape
        d_rs = torch.zeros(B, T, 2 * H, device=d_out.device, dtype=torch.float32)
        d_res = torch.zeros_like(d_rs)
        d_rs[..., H:].copy_(d_out.to(torch.float32).transpose(1, 2))
        d_rs.mul_((torch.arange(T, device=d_out.device)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(-1))   # * x_mask, :106
        ops.round_tf32_(d_rs.reshape(-1), d_rs.reshape(-1))
        dg, _ = self.backward_cl(d_rs, d_res, gW, _bias_grad_inline, _wgrad_inline, mask_input_grad=False)
        pk.unpack_grads()
        return d_res[..., :H].transpose(1, 2).contiguous(), (None if dg is None else dg.reshape(B, -1, 1))


# weight / bias gradient launches: on hifigan._Side's stream when it is on (inside a captured graph), else in line
def _bias_grad_inline(m, d, cols):
    if m.bias.grad is None:
        m.bias.grad = torch.zeros_like(m.bias)
    _Side.run(lambda: ops.colsum_(d.shape[0] * d.shape[1], cols, d.stride(1), d, m.bias.grad), d)


def _wgrad_inline(dy_, x_, shifts, out):
    _Side.run(lambda: ops.conv_wgrad(dy_, x_, shifts, out=out, accumulate=True), dy_, x_)


class PosteriorEncoder(nn.Module):
    """Drop-in for python/xvapitch/model.py:1422 ``PosteriorEncoder`` (configured at model.py:93-101: 513 -> 192 latent,
    WaveNet 192 x 16 layers, kernel 5, conditioning 512). ``forward(x [B, C, T], x_lengths, g, eps=None)`` returns
    (z, mean, log_scale, x_mask) in the reference's [B, C, T] layout; ``eps`` replays the N(0, 1) draw of model.py:1474
    (drawn with torch.randn on the device when omitted). ``backward(dz [B, C, T])`` accumulates the parameter gradients
    (the --hifi_only loss reaches the encoder through z only)."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, num_layers, cond_channels=0,
                 device=None, seed=1234):
        super().__init__()
        self.in_channels, self.out_channels, self.hidden_channels = in_channels, out_channels, hidden_channels
        self.kernel_size, self.dilation_rate, self.num_layers, self.cond_channels = kernel_size, dilation_rate, num_layers, cond_channels
        if out_channels % 32:
            raise NotImplementedError("out_channels must be a multiple of 32")
        self.in_cols = (in_channels + 31) // 32 * 32
        self.pre = _PlainConv(in_channels, hidden_channels, 1, bias_first=False)
        self.enc = WN(hidden_channels, hidden_channels, kernel_size, dilation_rate, num_layers, c_in_channels=cond_channels)
        self.proj = _PlainConv(hidden_channels, out_channels * 2, 1, bias_first=False)
        self.reset_parameters(seed)
        self.to(_need_cuda(device))
        self._packer = None
        self._ctx = None

    def reset_parameters(self, seed=1234):
        """torch's Conv1d default init (U(+-1/sqrt(fan_in)) for weights and biases), g = ||v|| as weight_norm sets it."""
        gen = torch.Generator().manual_seed(int(seed))
        with torch.no_grad():
            for m in self.modules():
                if not isinstance(m, (_WNConv, _PlainConv)):
                    continue
                v = m.weight if isinstance(m, _PlainConv) else m.weight_v
                bound = 1.0 / math.sqrt(v.shape[1] * v.shape[2])
                v.copy_((torch.rand(v.shape, generator=gen) * 2 - 1) * bound)
                if isinstance(m, _WNConv):
                    m.weight_g.copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1))
                m.bias.copy_((torch.rand(m.bias.shape, generator=gen) * 2 - 1) * bound)

    def _get_packer(self):
        if self._packer is None:
            pk = _WnPacker()
            # the 513-channel input rows are padded to 544 floats (whole 32-column chunks, 16-byte row pitch)
            pk.add_conv("pre", self.pre, self.pre.cout, self.pre.cin, 1, ld=self.in_cols)
            self.enc.register_weights(pk, "enc.")
            pk.add_conv("proj", self.proj, self.proj.cout, self.proj.cin, 1)
            pk.finalize(self.pre.weight.device)
            self._packer = pk
        return self._packer

    def forward(self, x, x_lengths, g=None, eps=None):
        B, cin, T = x.shape
        if cin != self.in_channels:
            raise ValueError(f"input has {cin} channels, the encoder was built for {self.in_channels}")
        dev = self.pre.weight.device
        keep = self.training
        lens = torch.as_tensor(x_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
        W = self._get_packer().pack()
        yp = torch.zeros(B, T, self.in_cols, device=dev, dtype=torch.float32)
        yp[..., :cin].copy_(x.to(torch.float32).transpose(1, 2))
        ops.round_tf32_(yp.reshape(-1), yp.reshape(-1))
        gr = None
        if g is not None and self.cond_channels > 0:
            gr = torch.empty(1, B, self.cond_channels, device=dev, dtype=torch.float32)
            ops.round_tf32_(g.to(torch.float32).reshape(-1).contiguous(), gr.reshape(-1))
        H = self.hidden_channels
        x0, x0r = torch.empty(B, T, H, device=dev, dtype=torch.float32), torch.empty(B, T, H, device=dev, dtype=torch.float32)
        ops.conv_fwd(yp, W["pre"][0], (0,), out=x0, bias=_bias(self.pre), lens=lens, out_act=x0r, out_act_slope=1.0)
        out, out_r = self.enc.forward_cl(x0, x0r, lens, gr, W, "enc.", keep=keep)
        stats = ops.conv_fwd(out_r, W["proj"][0], (0,), bias=_bias(self.proj), lens=lens)             # [B, T, 2C] masked
        if eps is None:
            eps = torch.randn(B, self.out_channels, T, device=dev, dtype=torch.float32)
        eps_cl = eps.to(device=dev, dtype=torch.float32).transpose(1, 2).contiguous()
        z = ops.vits_sample(stats, eps_cl, lens)                                                      # [B, T, C]
        self._ctx = (yp, out_r, stats, eps_cl, lens, W) if keep else None
        C = self.out_channels
        mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(1)
        self.z_cl = z                                  # channels-last view of z for the engine's own callers
        return z.transpose(1, 2), stats[..., :C].transpose(1, 2), stats[..., C:].transpose(1, 2), mask

    def backward(self, dz, channels_last=False):
        """dz: dL/dz, [B, C, T] (or [B, T, C] with channels_last)."""
        if self._ctx is None:
            raise RuntimeError("backward() needs a forward() in training mode first")
        yp, out_r, stats, eps_cl, lens, W = self._ctx
        pk = self._get_packer()
        gW = pk.zero_grads()
        B, T, _ = yp.shape
        H = self.hidden_channels
        dz_cl = (dz if channels_last else dz.transpose(1, 2)).to(torch.float32).contiguous()

        bias_grad, wgrad = _bias_grad_inline, _wgrad_inline
        dstats = ops.vits_sample_bwd(dz_cl, eps_cl, stats, lens)
        bias_grad(self.proj, dstats, self.proj.cout)
        wgrad(dstats, out_r, (0,), gW["proj"][0])
        d_rs = torch.zeros(B, T, 2 * H, device=yp.device, dtype=torch.float32)
        d_res = torch.zeros_like(d_rs)
        ops.conv_dgrad(dstats, W["proj"][0], (0,), out=d_rs[..., H:], lens=lens, round_out=True)       # dL/d(output), masked
        _, dx0 = self.enc.backward_cl(d_rs, d_res, gW, bias_grad, wgrad)
        bias_grad(self.pre, dx0, self.pre.cout)
        wgrad(dx0, yp, (0,), gW["pre"][0])
        pk.unpack_grads()
        self._ctx = None


class ResidualCouplingBlock(nn.Module):
    """Parameters of python/xvapitch/model.py:1476 ``ResidualCouplingBlock`` as xVAPitch builds it (mean_only, no
    projector): pre 1x1 (C/2 -> H), WN, post 1x1 (H -> C/2). Run by ResidualCouplingBlocks below."""

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, num_layers, dropout_p=0, cond_channels=0,
                 out_channels_override=None, mean_only=False):
        assert channels % 2 == 0, "channels should be divisible by 2"
        super().__init__()
        if out_channels_override or not mean_only:
            raise NotImplementedError("xVAPitch builds its flow with mean_only=True and no projector (model.py:1394-1403)")
        if (channels // 2) % 32:
            raise NotImplementedError("channels / 2 must be a multiple of 32")
        self.half_channels, self.mean_only, self.conv1d_projector = channels // 2, mean_only, None
        self.pre = _PlainConv(self.half_channels, hidden_channels, 1, bias_first=False)
        self.enc = WN(hidden_channels, hidden_channels, kernel_size, dilation_rate, num_layers, dropout_p=dropout_p,
                      c_in_channels=cond_channels)
        self.post = _PlainConv(hidden_channels, self.half_channels, 1, bias_first=False)


class ResidualCouplingBlocks(nn.Module):
    """Drop-in for python/xvapitch/model.py:1358 ``ResidualCouplingBlocks`` (configured at model.py:103-112: 192 channels,
    4 flows of a 4-layer WN, conditioning 512): ``forward(x [B, C, T], x_mask, g, reverse=False)`` -> [B, C, T]. Forward
    (posterior latent -> prior space) keeps what ``backward(dz)`` needs in training mode and returns (dL/dx, dL/dg);
    reverse is the inference / voice-conversion direction. The coupling x1 <- m + x1 * mask is the epilogue of the
    ``post`` GEMM (x1 in the residual slot, the mask in the lens slot); the channel flip between flows is a copy."""

    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, num_layers, num_flows=4, cond_channels=0,
                 args=None, device=None, seed=1234):
        super().__init__()
        if args is not None and getattr(args, "expanded_flow", False):
            raise NotImplementedError("expanded_flow")
        self.channels, self.hidden_channels, self.kernel_size = channels, hidden_channels, kernel_size
        self.dilation_rate, self.num_layers, self.num_flows, self.cond_channels = dilation_rate, num_layers, num_flows, cond_channels
        self.flows = nn.ModuleList([ResidualCouplingBlock(channels, hidden_channels, kernel_size, dilation_rate, num_layers,
                                                          cond_channels=cond_channels, mean_only=True) for _ in range(num_flows)])
        gen = torch.Generator().manual_seed(int(seed))
        with torch.no_grad():          # xavier_uniform_ on every weight (model.py:113-122), torch's default bias init
            for m in self.modules():
                if not isinstance(m, (_WNConv, _PlainConv)):
                    continue
                v = m.weight if isinstance(m, _PlainConv) else m.weight_v
                fan_in, fan_out = v.shape[1] * v.shape[2], v.shape[0] * v.shape[2]
                bound = math.sqrt(6.0 / (fan_in + fan_out))
                v.copy_((torch.rand(v.shape, generator=gen) * 2 - 1) * bound)
                if isinstance(m, _WNConv):
                    m.weight_g.copy_(v.flatten(1).norm(dim=1).view(-1, 1, 1))
                m.bias.copy_((torch.rand(m.bias.shape, generator=gen) * 2 - 1) / math.sqrt(fan_in))
        self.to(_need_cuda(device))
        self._packer = None
        self._ctx = None

    def _get_packer(self):
        if self._packer is None:
            pk = _WnPacker()
            for i, f in enumerate(self.flows):
                pk.add_conv(f"flows.{i}.pre", f.pre, f.pre.cout, f.pre.cin, 1)
                f.enc.register_weights(pk, f"flows.{i}.enc.")
                pk.add_conv(f"flows.{i}.post", f.post, f.post.cout, f.post.cin, 1)
            pk.finalize(self.flows[0].pre.weight.device)
            self._packer = pk
        return self._packer

    def forward(self, x, x_mask, g=None, reverse=False):
        B, C, T = x.shape
        dev = self.flows[0].pre.weight.device
        h = C // 2
        H = self.hidden_channels
        lens = x_mask.reshape(B, T).to(dev).sum(1).to(torch.int32)
        W = self._get_packer().pack()
        xc = x.to(device=dev, dtype=torch.float32).transpose(1, 2).contiguous()
        gr = None
        if g is not None and self.cond_channels > 0:
            gr = torch.empty(1, B, self.cond_channels, device=dev, dtype=torch.float32)
            ops.round_tf32_(g.to(device=dev, dtype=torch.float32).reshape(-1).contiguous(), gr.reshape(-1))
        keep = self.training and not reverse
        saved = []
        order = range(self.num_flows) if not reverse else reversed(range(self.num_flows))
        for i in order:
            f = self.flows[i]
            if reverse:
                xc = torch.flip(xc, [2])                                          # model.py:1419
            xr = torch.empty_like(xc)
            ops.round_tf32_(xc.reshape(-1), xr.reshape(-1))
            hx, hr = torch.empty(B, T, H, device=dev, dtype=torch.float32), torch.empty(B, T, H, device=dev, dtype=torch.float32)
            ops.conv_fwd(xr[..., :h], W[f"flows.{i}.pre"][0], (0,), out=hx, bias=_bias(f.pre), lens=lens, out_act=hr,
                         out_act_slope=1.0)                                        # pre(x0) * mask, :1530
            _, out_r = f.enc.forward_cl(hx, hr, lens, gr, W, f"flows.{i}.enc.", keep=keep)
            xn = torch.empty_like(xc)
            xn[..., :h].copy_(xc[..., :h])
            if not reverse:    # x1 <- m + x1 * mask = (post(h) + x1) * mask, :1539
                ops.conv_fwd(out_r, W[f"flows.{i}.post"][0], (0,), out=xn[..., h:], bias=_bias(f.post), residual=xc[..., h:],
                             lens=lens)
                xc = torch.flip(xn, [2])                                          # :1416
            else:              # x1 <- (x1 - m) * mask, :1544
                ops.conv_fwd(out_r, W[f"flows.{i}.post"][0], (0,), out=xn[..., h:], bias=-_bias(f.post), residual=xc[..., h:],
                             lens=lens, alpha=-1.0)
                xc = xn
            if keep:
                saved.append((xr, out_r))
        self._ctx = (saved, lens, gr, W) if keep else None
        return xc.transpose(1, 2)

    def backward(self, dz):
        """dz = dL/d(forward output) [B, C, T] -> (dL/dx [B, C, T], dL/dg [B, cond, 1] or None)."""
        if self._ctx is None:
            raise RuntimeError("backward() needs a forward() (not reverse) in training mode first")
        saved, lens, gr, W = self._ctx
        pk = self._get_packer()
        gW = pk.zero_grads()
        B, C, T = dz.shape
        h, H = C // 2, self.hidden_channels
        dev = dz.device
        d = dz.to(torch.float32).transpose(1, 2).contiguous()
        mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).to(torch.float32).unsqueeze(-1)
        dg = None
        for i in reversed(range(self.num_flows)):
            f = self.flows[i]
            xr, out_r = saved[i]
            d = torch.flip(d, [2])
            d1 = d[..., h:] * mask                                # gradient of both m (masked stats) and x1 * mask
            d1r = torch.empty_like(d1)
            ops.round_tf32_(d1.reshape(-1), d1r.reshape(-1))
            _bias_grad_inline(f.post, d1r, f.post.cout)
            _wgrad_inline(d1r, out_r, (0,), gW[f"flows.{i}.post"][0])
            d_rs = torch.zeros(B, T, 2 * H, device=dev, dtype=torch.float32)
            d_res = torch.zeros_like(d_rs)
            ops.conv_dgrad(d1r, W[f"flows.{i}.post"][0], (0,), out=d_rs[..., H:], lens=lens, round_out=True)
            dgi, dh = f.enc.backward_cl(d_rs, d_res, gW, _bias_grad_inline, _wgrad_inline, mask_input_grad=True)
            _bias_grad_inline(f.pre, dh, f.pre.cout)
            _wgrad_inline(dh, xr[..., :h], (0,), gW[f"flows.{i}.pre"][0])
            dn = torch.empty_like(d)
            ops.conv_dgrad(dh, W[f"flows.{i}.pre"][0], (0,), out=dn[..., :h], residual=d[..., :h])
            dn[..., h:].copy_(d1)
            d = dn
            if dgi is not None:
                dg = dgi if dg is None else dg + dgi
        pk.unpack_grads()
        self._ctx = None
        return d.transpose(1, 2).contiguous(), (None if dg is None else dg.reshape(B, -1, 1))


def prior_alignment(z_p, m_p, logs_p, x_lengths, y_lengths):
    """The alignment block of xVAPitch.train_step (python/xvapitch/model.py:763-777, 855-856) on the device: prior
    log-likelihood of every latent frame under every text token (one exact-fp32 batched product), the monotonic
    alignment search (``xva_mas_width1``; the reference copies the [B, t_text, t_spec] matrix to the host and runs a numpy
    loop, every step), the durations and the prior expanded to the frames (a gather by the path, not the reference's
    one-hot einsum). Reference layouts in and out: z_p [B, C, Ts], m_p / logs_p [B, C, Tt] ->
    dict(attn [B, 1, Tt, Ts], durations [B, 1, Tt], m_p [B, C, Ts], logs_p [B, C, Ts], cum) -- ``cum`` (int32 prefix sums
    of the durations) is what ``prior_expand_backward`` needs."""
    dev = z_p.device
    B, C, Ts = z_p.shape
    Tt = m_p.shape[2]
    xl = torch.as_tensor(x_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
    yl = torch.as_tensor(y_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
    zc = z_p.detach().to(torch.float32).transpose(1, 2).contiguous()
    mc = m_p.to(torch.float32).transpose(1, 2).contiguous()
    lc = logs_p.to(torch.float32).transpose(1, 2).contiguous()
    logp = ops.vits_logp(zc, mc.detach(), lc.detach())                                   # [B, Ts, Tt]
    hard, durs = ops.mas_width1(logp, xl, yl, is_log=True, stay_on_tie=True)             # [B, Ts, Tt] 0 / 1, [B, Tt]
    cum, _ = ops.duration_scan(durs.to(torch.float32), 1.0, Ts)
    m_e = ops.regulate_gather(mc, cum, Ts)
    l_e = ops.regulate_gather(lc, cum, Ts)
    return {"attn": hard.transpose(1, 2).unsqueeze(1), "durations": durs.to(torch.float32).unsqueeze(1), "logp": logp,
            "m_p": m_e.transpose(1, 2), "logs_p": l_e.transpose(1, 2), "cum": cum}


def prior_expand_backward(d_m_p, d_logs_p, cum, t_text):
    """Gradients of the expanded prior [B, C, Ts] back to the per-token prior [B, C, Tt] (the path is a constant)."""
    dm = ops.regulate_scatter(d_m_p.transpose(1, 2).contiguous(), cum, t_text)
    dl = ops.regulate_scatter(d_logs_p.transpose(1, 2).contiguous(), cum, t_text)
    return dm.transpose(1, 2), dl.transpose(1, 2)


def kl_loss(z_p, logs_q, m_p, logs_p, y_lengths, scale=1.0):
    """VitsGeneratorLoss.kl_loss (python/xvapitch/losses.py:86-103) with a prefix mask given as lengths, and its gradients:
    [B, C, T] tensors in -> (loss, (dz_p, dlogs_q, dm_p, dlogs_p) [B, C, T], already times ``scale``)."""
    dev = z_p.device
    yl = torch.as_tensor(y_lengths).reshape(-1).to(device=dev, dtype=torch.int32)
    cl = lambda t: t.detach().to(torch.float32).transpose(1, 2).contiguous()
    loss, grads = ops.vits_kl(cl(z_p), cl(logs_q), cl(m_p), cl(logs_p), yl, scale)
    return loss, tuple(g.transpose(1, 2) for g in grads)


class HifiOnlyStep:
    """One xVAPitch ``--hifi_only`` iteration (amp off, gam 1): posterior encoder -> random 32-frame latent segment ->
    waveform decoder -> VITS discriminator; generator-side loss = 45 * L1(log-mel) + LSGAN (the feature-matching term
    is reported but, as the reference calls it, carries no gradient to the generator); discriminator loss on the same
    scores; AdamW(lr, betas (0.8, 0.99), eps 1e-9, weight decay 0.01) on encoder + decoder and AdamW(2e-4, ...) on the
    discriminator, both stepped after both backward passes (xva_train.py:651-736). The reference runs the discriminator
    three times per iteration (generator pass, and real + fake again for its own loss); its weights do not change in
    between, so ONE batched pass over (real, fake) serves both backward passes here.

    ``step(linear [B, 513, T], y_lengths, waveform [B, 1, 256 T], d_vectors [B, 512], eps=None, u=None)``: eps / u replay
    the posterior sample and the segment draw (model.py:1474, util.py:162). Returns the reference's loss dict."""

    def __init__(self, posterior_encoder, waveform_decoder, disc, lr=0.000175, disc_lr=0.0002, betas=(0.8, 0.99), eps=1e-9,
                 weight_decay=0.01, segment=SPEC_SEGMENT, world=1, group=None):
        """world > 1: one process per GPU, each on its own utterances; the two gradient arenas are all-reduced (mean)
        right before their optimizer steps -- the reference's nn.DataParallel semantics, whose criterion runs inside the
        replicated forward and whose trainer averages the per-replica losses (xvapitch/xva_train.py:678-683)."""
        self.enc, self.dec, self.disc, self.segment = posterior_encoder, waveform_decoder, disc, int(segment)
        self.world, self.group = int(world), group
        dev = next(waveform_decoder.parameters()).device
        self.mel = MelSpectrogram.vits(device=dev)
        self.optim_g = AdamW(list(posterior_encoder.parameters()) + list(waveform_decoder.parameters()), lr, betas, eps,
                             weight_decay)                                                  # training_util.py:31-32
        self.optim_d = AdamW(disc.parameters(), disc_lr, betas, eps, weight_decay)
        self.steps = 0

    def _all_reduce(self, flat_grad):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            flat_grad.mul_(1.0 / self.world)

    def step(self, linear, y_lengths, waveform, d_vectors, eps=None, u=None):
        """No host synchronisation when y_lengths (int32) and u are device tensors: the segment starts stay on the device
        and the segments are gathered / scattered by index, so the whole iteration can be captured in a CUDA graph."""
        enc, dec, disc, S = self.enc, self.dec, self.disc, self.segment
        dev = next(dec.parameters()).device
        B, _, T = linear.shape
        if not (torch.is_tensor(y_lengths) and y_lengths.is_cuda):
            lens_host = [int(v) for v in torch.as_tensor(y_lengths).reshape(-1).tolist()]
            if min(lens_host) - S + 1 <= 0:
                raise ValueError(" [!] At least one sample is shorter than the segment size.")      # util.py:161
            y_lengths = torch.tensor(lens_host, dtype=torch.int32, device=dev)
        lens = y_lengths.reshape(-1).to(torch.int32)
        self.optim_g.zero_grad()
        self.optim_d.zero_grad()
        g = torch.nn.functional.normalize(d_vectors.to(device=dev, dtype=torch.float32)).unsqueeze(-1)    # model.py:920
        enc(linear.to(dev), lens, g=g, eps=eps)
        z = enc.z_cl                                                                         # [B, T, C]
        if u is None:
            u = torch.rand(B, device=dev)
        starts = (u.to(device=dev, dtype=torch.float32) * (lens - S + 1)).long()             # util.py:160-162
        rows = torch.arange(B, device=dev)[:, None]
        fidx = starts[:, None] + torch.arange(S, device=dev)[None, :]                        # frames of each segment
        z_seg = z[rows, fidx]                                                                # [B, S, C]
        o = dec(z_seg.transpose(1, 2), g=g)                                                  # [B, 1, S * 256]
        wav = waveform.to(device=dev, dtype=torch.float32).reshape(B, -1)
        wav_seg = wav.gather(1, starts[:, None] * HOP + torch.arange(S * HOP, device=dev)[None, :])    # util.py:165-178
        fake = o.reshape(B, -1)
        # ---- one discriminator pass over (real, fake); discriminator loss and its parameter gradients
        rs, frs, gs, fgs = disc(wav_seg, fake)
        loss_disc = discriminator_loss_backward(disc, rs, gs)
        # ---- generator side: mel + adversarial gradient wrt the generated waveform
        mel_real = self.mel(wav_seg)
        mel_fake = self.mel(fake)
        n = mel_fake.numel()
        acc = torch.zeros(1, device=dev, dtype=torch.float64)
        ops.reduce_l1(mel_real, mel_fake, acc)
        loss_mel = 45.0 * acc[0] / n
        dwave = self.mel.backward(ops.l1_grad(mel_real, mel_fake, 45.0 / n))
        loss_gen, loss_feat = generator_adv_loss_backward(disc, gs, frs, fgs, dwave, pools=0, fm_grad=False)
        dz_seg, _ = dec.backward(dwave.view(B, 1, -1), need_input_grad=True)                 # [B, C, S]
        dz = torch.zeros_like(z)
        dz[rows, fidx] = dz_seg.transpose(1, 2)
        enc.backward(dz, channels_last=True)
        self._all_reduce(self.optim_d.g)
        self._all_reduce(self.optim_g.g)
        self.optim_g.step()
        self.optim_d.step()
        self.steps += 1
        return {"loss": loss_feat + loss_mel + loss_gen, "loss_gen": loss_gen, "loss_feat": loss_feat, "loss_mel": loss_mel,
                "loss_disc": loss_disc, "slice_ids": starts}
