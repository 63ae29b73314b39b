"""Tensor-level wrappers over the C ABI: torch tensors in, torch tensors out, kernels from libxva_b200.so.

torch is used for device memory and the current stream only. All tensors are fp32, channels-last [batch, time,
channels] with a contiguous last dimension (row / batch strides are passed through, so slices such as the q, k, v
thirds of a fused qkv projection are used in place).
"""
import ctypes as C

import torch

from . import capi


def capturing():
    """True while the current CUDA stream is being captured into a graph (False on a host without a CUDA context: the
    launch-sequence dry run of tests/)."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class nvtx:
    """NVTX range around one stage of a step (forward / loss / backward sections show up by name in nsys / ncu timelines).
    XVA_NVTX=0 turns the ranges off; they cost ~1 us of host time each and nothing on the device."""
    _on = None

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if nvtx._on is None:
            import os
            nvtx._on = os.environ.get("XVA_NVTX", "1") != "0" and torch.cuda.is_available()
        if nvtx._on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if nvtx._on:
            torch.cuda.nvtx.range_pop()
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check3(t, name):
    if t.dtype != torch.float32 or t.dim() != 3 or t.stride(2) != 1 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA fp32 [batch, rows, cols] tensor with contiguous last dim, got "
                         f"{t.dtype} {tuple(t.shape)} strides {t.stride()}")


def gemm_launch(args, ref=False):
    """Launch one tap-GEMM described by a filled capi.GemmArgs."""
    capi.call("xva_gemm_ref" if ref else "xva_gemm", C.byref(args), _stream())


def _base_args(mode, shifts):
    g = capi.GemmArgs()
    g.mode = mode
    g.taps = len(shifts)
    for j, s in enumerate(shifts):
        g.shift[j] = int(s)
    g.ZR = 1
    g.split = 1
    g.b_nz = 1
    g.alpha = 1.0
    g.ln_eps = 1e-5
    return g


def _epilogue(g, out, bias=None, relu=False, gate=None, gate_slope=0.0, residual=None, lens=None, ln=None,
              ln_eps=1e-5, save_ln=False, drop_p=0.0, drop_post=False, seed=0, seed_dev=None, alpha=1.0,
              round_out=False, act_slope=0.0, out_act=None, out_act_slope=0.0, tanh=False, halo=False):
    keep = [out, bias, gate, residual, lens]
    g.out, g.o_rs, g.o_zs = _p(out), out.stride(1), out.stride(0)
    g.alpha = alpha
    flags = 0
    if bias is not None:
        g.bias = _p(bias)
    if relu or act_slope != 0.0:
        flags |= capi.GEMM_RELU
        g.act_slope = act_slope
    if tanh:
        flags |= capi.GEMM_TANH
    if out_act is not None:
        assert out_act.stride() == out.stride()
        g.out_act, g.out_act_slope = _p(out_act), out_act_slope
        keep.append(out_act)
    if gate is not None:
        _check3(gate, "gate")
        g.gate, g.g_rs, g.g_zs, g.gate_slope = _p(gate), gate.stride(1), gate.stride(0), gate_slope
    if residual is not None:
        _check3(residual, "residual")
        g.residual, g.r_rs, g.r_zs = _p(residual), residual.stride(1), residual.stride(0)
    if lens is not None:
        assert lens.dtype == torch.int32
        g.lens = _p(lens)
    extra = {}
    if ln is not None:
        gamma, beta = ln
        flags |= capi.GEMM_LN
        g.gamma, g.beta, g.ln_eps = _p(gamma), _p(beta), ln_eps
        keep += [gamma, beta]
        if save_ln:
            pre = torch.empty_like(out)
            mean = torch.empty(out.shape[0] * out.shape[1], device=out.device, dtype=torch.float32)
            rstd = torch.empty_like(mean)
            assert pre.stride() == out.stride()
            g.out_pre, g.ln_mean, g.ln_rstd = _p(pre), _p(mean), _p(rstd)
            extra = {"pre": pre, "mean": mean, "rstd": rstd}
    if drop_p > 0.0:
        flags |= capi.GEMM_DROP_POST if drop_post else capi.GEMM_DROP_PRE
        g.drop_p, g.seed, g.seed_dev = drop_p, seed, _p(seed_dev)
    if round_out:
        flags |= capi.GEMM_ROUND_OUT
    if halo:
        flags |= capi.GEMM_HALO
    g.flags = flags
    return keep, extra


def conv_fwd(x, wp, shifts=(0,), out=None, ref=False, a_cols=None, out_rows=None, groups=1, grp_step=0, **epi):
    """out[b,t,n] = sum_j sum_k x[b, t+shifts[j], a_cols[j]+k] * wp[j, n, k]  (+ epilogue).
    x [B,T,Kx], wp [taps,N,K]; a_cols (per-tap column offsets into x's rows, default 0) lets a strided convolution
    run on the [T/stride, stride*C] view of its input. groups > 1: grouped convolution in one launch -- output columns
    [g*N/groups, (g+1)*N/groups) read x columns a_cols[j] + g*grp_step + [0, K) (xva_gemm_args.groups)."""
    _check3(x, "x")
    _check3(wp, "wp")
    B, T, Kx = x.shape
    taps, N, K = wp.shape
    assert taps == len(shifts) and (a_cols is not None or groups > 1 or Kx == K)
    R = T if out_rows is None else int(out_rows)   # output rows per item; x keeps its own row count (a_rows)
    if out is None:
        out = torch.empty(B, R, N, device=x.device, dtype=torch.float32)
    g = _base_args(0, shifts)
    if a_cols is not None:
        for j, c in enumerate(a_cols):
            assert c + (groups - 1) * grp_step + K <= Kx
            g.a_col[j] = int(c)
    g.Z, g.R, g.N, g.K = B, R, N, K
    g.groups, g.grp_step = int(groups), int(grp_step)
    g.a, g.a_rs, g.a_zs, g.a_rows = _p(x), x.stride(1), x.stride(0), T
    g.b, g.b_rs, g.b_zs, g.b_nz, g.b_tap_z = _p(wp), wp.stride(1), wp.stride(0), taps, 1
    keep, extra = _epilogue(g, out, **epi)
    gemm_launch(g, ref)
    return (out, extra) if extra else out


def conv_dgrad(dy, wp, shifts=(0,), out=None, ref=False, out_rows=None, groups=1, **epi):
    """dx[b,t,k] = sum_j sum_n dy[b, t-shifts[j], n] * wp[j, n, k]: input gradient of conv_fwd, same packed weights
    read MN-major (no transposed copy).  dy [B,T,N], wp [taps,N,K]. groups > 1: wp's rows [g*N/groups, ...) are group
    g's filters and its K columns the group's input channels; dx has groups*K columns."""
    _check3(dy, "dy")
    _check3(wp, "wp")
    B, T, N = dy.shape
    taps, N2, K = wp.shape
    assert N2 == N and taps == len(shifts)
    R = T if out_rows is None else int(out_rows)
    if out is None:
        out = torch.empty(B, R, K * groups, device=dy.device, dtype=torch.float32)
    g = _base_args(1, [-s for s in shifts])
    g.Z, g.R, g.N, g.K = B, R, K * groups, N // groups
    g.groups = int(groups)
    g.a, g.a_rs, g.a_zs, g.a_rows = _p(dy), dy.stride(1), dy.stride(0), T
    g.b, g.b_rs, g.b_zs, g.b_nz, g.b_tap_z = _p(wp), wp.stride(1), wp.stride(0), taps, 1
    keep, extra = _epilogue(g, out, **epi)
    gemm_launch(g, ref)
    return (out, extra) if extra else out


def conv_wgrad(dy, x, shifts=(0,), out=None, accumulate=False, split=None, ref=False, x_cols=None, n_cols=None,
               dy_rows=None, groups=1, grp_step=0):
    """dw[j,n,k] (+)= sum_b sum_t dy[b,t,n] * x[b, t+shifts[j], k]: weight gradient of conv_fwd in the packed layout.
    dy [B,T,N], x [B,T,K] -> dw [taps,N,K].  N and K must be multiples of 32."""
    _check3(dy, "dy")
    _check3(x, "x")
    B, T, N = dy.shape
    B2, T2, K = x.shape
    assert B2 == B
    taps = len(shifts)
    if n_cols is not None:      # x_cols[j] + [0, n_cols) are the columns of x tap j contracts with
        K = int(n_cols)
    if out is None:
        out = torch.zeros(taps, N, K, device=dy.device, dtype=torch.float32)
        accumulate = True
    g = _base_args(2, shifts)
    if x_cols is not None:
        for j, c in enumerate(x_cols):
            g.a_col[j] = int(c)
    g.Z, g.R, g.M, g.N, g.ZR = B, T, N, K, B
    g.a, g.a_rs, g.a_zs, g.a_rows = _p(dy), dy.stride(1), dy.stride(0), (T if dy_rows is None else int(dy_rows))
    g.b, g.b_rs, g.b_zs, g.b_rows = _p(x), x.stride(1), x.stride(0), T2
    g.out, g.o_rs, g.o_zs, g.o_js = _p(out), out.stride(1), 0, out.stride(0)
    g.groups, g.grp_step = int(groups), int(grp_step)
    if split is None:
        tiles = taps * ((N + 127) // 128) * ((K + 255) // 256)
        # CTAs per output tile: up to two waves in total, at most one per item. (The kernel can also cut every item's
        # rows into chunks -- pass split > B -- but measured on the HiFi-GAN generator that is slower: the per-tap tiles
        # of a small-channel weight gradient are HBM-bound on re-reading dy / x once per tap, not short of CTAs.)
        split = max(1, min(B, (2 * 148) // max(tiles, 1)))
    g.split = split if accumulate else 1
    g.flags = capi.GEMM_ATOMIC if accumulate else 0
    gemm_launch(g, ref)
    return out


def bmm_nt(a, b, alpha=1.0, out=None, ref=False, round_out=False, softmax_bwd=None):
    """out[z,m,n] = alpha * sum_k a[z,m,k] * b[z,n,k]   (torch.bmm(a, b.transpose(1,2))).
    softmax_bwd = (P [Z,M,ld], rowvec [Z*M], drop_p, seed, seed_dev): the product is dP_dropped = dO.V^T and the epilogue
    turns it into alpha * dS = alpha * P * (dP_dropped * mask - rowvec) (XVA_GEMM_SOFTMAX_BWD); out's pad columns
    [N, ld) must already be zero."""
    _check3(a, "a")
    _check3(b, "b")
    Z, M, K = a.shape
    N = b.shape[1]
    if out is None:
        out = torch.empty(Z, M, N, device=a.device, dtype=torch.float32)
    g = _base_args(0, (0,))
    g.Z, g.R, g.N, g.K = Z, M, N, K
    g.a, g.a_rs, g.a_zs = _p(a), a.stride(1), a.stride(0)
    g.b, g.b_rs, g.b_zs, g.b_nz, g.b_batch_z = _p(b), b.stride(1), b.stride(0), Z, 1
    if softmax_bwd is None:
        _epilogue(g, out, alpha=alpha, round_out=round_out)
    else:
        P, rowvec, drop_p, seed, seed_dev = softmax_bwd
        _epilogue(g, out, alpha=alpha, round_out=round_out, gate=P[..., :N])
        g.flags |= capi.GEMM_SOFTMAX_BWD
        g.rowvec, g.drop_ld = _p(rowvec), P.stride(1)
        g.drop_p, g.seed, g.seed_dev = float(drop_p), int(seed), _p(seed_dev)
    gemm_launch(g, ref)
    return out


def rowdot2(a, b):
    """[Z,R,C] x [Z,R,C] -> [Z*R] row-wise dot products (contiguous last dim; row pitches are passed through)."""
    _check3(a, "a")
    _check3(b, "b")
    Z, R, Cc = a.shape
    assert a.stride(0) == R * a.stride(1) and b.stride(0) == R * b.stride(1)
    out = torch.empty(Z * R, device=a.device, dtype=torch.float32)
    capi.call("xva_rowdot2", _p(a), _p(b), Z * R, Cc, a.stride(1), b.stride(1), _p(out), _stream())
    return out


def bmm_nn(a, b, alpha=1.0, out=None, ref=False, round_out=False):
    """out[z,m,n] = alpha * sum_k a[z,m,k] * b[z,k,n]   (torch.bmm(a, b)); n must be a multiple of 32."""
    _check3(a, "a")
    _check3(b, "b")
    Z, M, K = a.shape
    N = b.shape[2]
    if out is None:
        out = torch.empty(Z, M, N, device=a.device, dtype=torch.float32)
    g = _base_args(1, (0,))
    g.Z, g.R, g.N, g.K = Z, M, N, K
    g.a, g.a_rs, g.a_zs = _p(a), a.stride(1), a.stride(0)
    g.b, g.b_rs, g.b_zs, g.b_nz, g.b_batch_z = _p(b), b.stride(1), b.stride(0), Z, 1
    _epilogue(g, out, alpha=alpha, round_out=round_out)
    gemm_launch(g, ref)
    return out


def bmm_tn(a, b, alpha=1.0, out=None, ref=False, round_out=False):
    """out[z,m,n] = alpha * sum_t a[z,t,m] * b[z,t,n]   (torch.bmm(a.transpose(1,2), b)); m, n multiples of 32."""
    _check3(a, "a")
    _check3(b, "b")
    Z, T, M = a.shape
    N = b.shape[2]
    if out is None:
        out = torch.empty(Z, M, N, device=a.device, dtype=torch.float32)
    g = _base_args(2, (0,))
    g.Z, g.R, g.M, g.N, g.ZR = Z, T, M, N, 1
    g.a, g.a_rs, g.a_zs, g.a_rows = _p(a), a.stride(1), a.stride(0), T
    g.b, g.b_rs, g.b_zs, g.b_rows = _p(b), b.stride(1), b.stride(0), T
    g.out, g.o_rs, g.o_zs, g.o_js = _p(out), out.stride(1), out.stride(0), 0
    g.alpha = alpha
    g.flags = capi.GEMM_ROUND_OUT if round_out else 0
    gemm_launch(g, ref)
    return out


# ---------------------------------------------------------------------------------------------- length regulator
def duration_scan(durs, pace=1.0, mel_max_len=None):
    """-> (cum int32 [B,Tt+1], dec_lens int32 [B])   fastpitch/model.py:62-65,76-78."""
    B, Tt = durs.shape
    d = durs.float().contiguous()
    cum = torch.empty(B, Tt + 1, device=d.device, dtype=torch.int32)
    dec = torch.empty(B, device=d.device, dtype=torch.int32)
    capi.call("xva_regulate_len_scan", _p(d), B, Tt, float(pace), -1 if mel_max_len is None else int(mel_max_len),
              _p(cum), _p(dec), _stream())
    return cum, dec


def regulate_gather(enc, cum, T_out, want_idx=False):
    B, Tt, Cc = enc.shape
    enc = enc.contiguous()
    out = torch.empty(B, T_out, Cc, device=enc.device, dtype=torch.float32)
    idx = torch.empty(B, T_out, device=enc.device, dtype=torch.int32) if want_idx else None
    capi.call("xva_regulate_len_fwd", _p(enc), _p(cum), B, Tt, Cc, T_out, _p(out), _p(idx), _stream())
    return (out, idx) if want_idx else out


def regulate_scatter(dout, cum, Tt, out=None):
    B, T_out, Cc = dout.shape
    dout = dout.contiguous()
    acc = out is not None
    if out is None:
        out = torch.empty(B, Tt, Cc, device=dout.device, dtype=torch.float32)
    capi.call("xva_regulate_len_bwd", _p(dout), _p(cum), B, Tt, Cc, T_out, _p(out), int(acc), _stream())
    return out


def average_pitch(pitch, durs, log1p=False):
    """fastpitch/model.py:82-100.  pitch [B,F,Tm], durs [B,Tt] -> [B,F,Tt]."""
    B, F, Tm = pitch.shape
    Tt = durs.shape[1]
    pitch = pitch.float().contiguous()
    d = durs.float().contiguous()
    out = torch.empty(B, F, Tt, device=pitch.device, dtype=torch.float32)
    capi.call("xva_average_pitch", _p(pitch), _p(d), B, F, Tm, Tt, _p(out), int(log1p), _stream())
    return out


def mas_width1(attn, in_lens, out_lens, is_log=False, stay_on_tie=False):
    """b_mas, fastpitch/alignment.py:110-118. attn [B, 1, Tm, Tt] (or [B, Tm, Tt]) soft alignment, in_lens / out_lens [B]
    -> (attn_hard, same shape, 0/1 floats; durations int32 [B, Tt] = attn_hard.sum over mel). stay_on_tie (with is_log):
    xVAPitch's maximum_path (xvapitch/util.py:14-53) on [B, t_mel, t_text] log-likelihoods -- the reference's
    value.transpose(1, 2); see maximum_path below."""
    shape = attn.shape
    a = attn.reshape(shape[0], shape[-2], shape[-1]).to(torch.float32).contiguous()
    B, Tm, Tt = a.shape
    hard = torch.empty_like(a)
    durs = torch.empty(B, Tt, device=a.device, dtype=torch.int32)
    il, ol = in_lens.to(torch.int32).contiguous(), out_lens.to(torch.int32).contiguous()   # keep both alive for the call
    if not is_log:
        # the logarithm the search takes of every probability, as one parallel pass (same function, same result) instead
        # of a double-precision log inside each of the Tm sequential steps
        la = torch.empty_like(a)
        capi.call("xva_mas_log", _p(a), a.numel(), _p(la), _stream())
        a = la
    capi.call("xva_mas_width1", _p(a), _p(il), _p(ol), B, Tm, Tt, 1 | (2 if stay_on_tie else 0), _p(hard), _p(durs), _stream())
    return hard.view(shape), durs


def maximum_path(value, x_lens, y_lens):
    """xVAPitch's monotonic alignment search, python/xvapitch/util.py:14-53, on the device (the reference copies `value` to
    the host and runs a numpy loop at every training step, xvapitch/model.py:776): value [B, t_text, t_mel]
    log-likelihoods, mask = (text < x_lens) & (mel < y_lens) -> path [B, t_text, t_mel] of 0 / 1."""
    hard, _ = mas_width1(value.transpose(1, 2), x_lens, y_lens, is_log=True, stay_on_tie=True)
    return hard.transpose(1, 2)


def attn_score_fwd(q, k, prior, in_lens):
    """ConvAttention.forward after the projections, fastpitch/attention.py:203-219. q [B,Tm,C], k [B,Tt,C] (row pitches
    passed through), prior [B,Tm,Tt], in_lens int32 [B] -> (attn_logprob, attn_soft, prior as the contiguous fp32 tensor
    the kernel read -- the backward needs the same values), all [B,Tm,Tt]."""
    _check3(q, "q")
    _check3(k, "k")
    B, Tm, Cc = q.shape
    Tt = k.shape[1]
    assert q.stride(0) == Tm * q.stride(1) and k.stride(0) == Tt * k.stride(1) and in_lens.dtype == torch.int32
    pr = prior.to(torch.float32).contiguous()
    assert tuple(pr.shape) == (B, Tm, Tt), (tuple(pr.shape), (B, Tm, Tt))
    logprob = torch.empty(B, Tm, Tt, device=q.device, dtype=torch.float32)
    soft = torch.empty_like(logprob)
    capi.call("xva_attn_score_fwd", _p(q), q.stride(1), _p(k), k.stride(1), _p(pr), _p(in_lens), B, Tm, Tt, Cc,
              _p(logprob), _p(soft), _stream())
    return logprob, soft, pr


def attn_score_bwd(g, logprob, prior, q, k, ld=96):
    """-> (dq [B,Tm,C], dk [B,Tt,C]) as views of zero-tailed buffers with row pitch ``ld`` (they are the MN-major
    operands of the projection stacks' weight-gradient GEMMs). g [B,Tm,Tt] is overwritten with the raw-score gradient."""
    B, Tm, Cc = q.shape
    Tt = k.shape[1]
    dq = torch.zeros(B, Tm, ld, device=q.device, dtype=torch.float32)
    dk = torch.zeros(B, Tt, ld, device=q.device, dtype=torch.float32)
    capi.call("xva_attn_score_bwd", _p(g), _p(logprob), _p(prior), _p(q), q.stride(1), _p(k), k.stride(1), B, Tm, Tt, Cc,
              _p(g), _p(dq), ld, _p(dk), ld, _stream())
    return dq[..., :Cc], dk[..., :Cc]


def attn_ctc(logprob, in_lens, out_lens, blank_logprob=-1.0):
    """AttentionCTCLoss, fastpitch/attn_loss_function.py:20-44, whole batch in one launch.
    -> (cost double [B] per utterance, grad [B,Tm,Tt] = d(mean cost)/d(logprob))."""
    B, Tm, Tt = logprob.shape
    assert logprob.is_contiguous() and in_lens.dtype == torch.int32 and out_lens.dtype == torch.int32
    nbytes = int(capi.load().xva_attn_ctc_workspace_bytes(B, Tm, Tt))
    ws = torch.empty(nbytes // 8, device=logprob.device, dtype=torch.float64)
    cost = torch.empty(B, device=logprob.device, dtype=torch.float64)
    grad = torch.empty_like(logprob)
    capi.call("xva_attn_ctc", _p(logprob), _p(in_lens), _p(out_lens), B, Tm, Tt, float(blank_logprob), _p(ws), nbytes,
              _p(cost), _p(grad), _stream())
    return cost, grad


def attn_bin_loss(hard, soft, acc, eps=1e-12):
    """acc (double[2]) += {sum_{hard==1} log(max(soft, eps)), sum(hard)}  (attn_loss_function.py:47-54)."""
    assert hard.is_contiguous() and soft.is_contiguous() and hard.shape == soft.shape and acc.dtype == torch.float64
    Tt = hard.shape[-1]
    capi.call("xva_attn_bin_loss", _p(hard), _p(soft), hard.numel() // Tt, Tt, float(eps), _p(acc), _stream())


def attn_grad_combine(gctc, a, hard=None, soft=None, acc=None, bw=0.0, eps=1e-12):
    """g = a * gctc + (bw / acc[1]) * (soft * rowsum(h') - h'), h' = hard * [soft >= eps]; [.., Tt] tensors."""
    Tt = gctc.shape[-1]
    g = torch.empty_like(gctc)
    use_kl = hard is not None and bw != 0.0
    capi.call("xva_attn_grad_combine", _p(gctc), _p(hard) if use_kl else None, _p(soft) if use_kl else None,
              _p(acc) if use_kl else None, float(a), float(bw), float(eps), gctc.numel() // Tt, Tt, _p(g), _stream())
    return g


def attn_fwd(qkv, lens, scale, drop_p=0.0, seed=0, seed_dev=None, drop_ld=None):
    """Fused attention of one FFT block (xva_attn_fwd): qkv [B, T, 192] (q | k | v, tf32-rounded) -> (vec [B, T, 64]
    tf32-rounded, lse [B, T]). lens int32 [B] masks the keys >= lens[b]."""
    _check3(qkv, "qkv")
    B, T, C = qkv.shape
    assert C == 192 and (lens is None or lens.dtype == torch.int32)
    out = torch.empty(B, T, 64, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B, T, device=qkv.device, dtype=torch.float32)
    capi.call("xva_attn_fwd", _p(qkv), qkv.stride(1), qkv.stride(0), B, T, _p(lens), float(scale), float(drop_p),
              int(seed) & 0xFFFFFFFFFFFFFFFF, _p(seed_dev), int(drop_ld if drop_ld else (T + 31) // 32 * 32), _p(out),
              out.stride(1), out.stride(0), _p(lse), _stream())
    return out, lse


def attn_bwd(qkv, dvec, vec, lse, lens, scale, drop_p=0.0, seed=0, seed_dev=None, drop_ld=None, out=None):
    """Backward of attn_fwd (xva_attn_bwd): -> dqkv [B, T, 192] = dq | dk | dv, tf32-rounded. dvec = gradient of the
    attention output (tf32-rounded), vec / lse = what attn_fwd returned."""
    _check3(qkv, "qkv")
    _check3(dvec, "dvec")
    B, T, _ = qkv.shape
    dsum = rowdot2(dvec, vec)
    dqkv = torch.empty_like(qkv) if out is None else out
    capi.call("xva_attn_bwd", _p(qkv), qkv.stride(1), qkv.stride(0), _p(dvec), dvec.stride(1), dvec.stride(0), _p(lse),
              _p(dsum), B, T, _p(lens), float(scale), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(seed_dev),
              int(drop_ld if drop_ld else (T + 31) // 32 * 32), _p(dqkv), dqkv.stride(1), dqkv.stride(0), _stream())
    return dqkv


# ---------------------------------------------------------------------------------------------- row kernels
def softmax_fwd(s, lens, n_valid, drop_p=0.0, seed=0, seed_dev=None):
    """transformer.py:120-127.  s [Z,R,ld] holds alpha*q.k^T in its first n_valid columns -> (p, pd); pd is p when
    drop_p == 0. Pad columns [n_valid, ld) of the outputs are zero."""
    Z, R, ld = s.shape
    p = torch.empty_like(s)
    pd = torch.empty_like(s) if drop_p > 0.0 else None
    capi.call("xva_softmax_fwd", _p(s), _p(lens), Z, R, int(n_valid), ld, _p(p), _p(pd), float(drop_p), int(seed),
              _p(seed_dev), _stream())
    return p, (pd if pd is not None else p)


def softmax_bwd_(p, dpd, n_valid, alpha, drop_p=0.0, seed=0, seed_dev=None):
    """In place: dpd (gradient wrt the dropped-out probabilities) becomes alpha * d(scores)."""
    Z, R, ld = p.shape
    capi.call("xva_softmax_bwd", _p(p), _p(dpd), Z, R, int(n_valid), ld, float(alpha), float(drop_p), int(seed),
              _p(seed_dev), _stream())
    return dpd


def layernorm_bwd(dy, saved, gamma, lens, dgamma, dbeta, dbias=None, want_drop=False, drop_post_p=0.0, seed_post=0,
                  drop_pre_p=0.0, seed_pre=0, seed_dev=None, relu_gate=False):
    """Backward of the LayerNorm epilogue. saved = {"pre","mean","rstd"} from conv_fwd(save_ln=True).
    -> dx (gradient wrt the pre-LN sum), dx_drop (dx * pre-dropout mask, or dx itself when no dropout)."""
    Z, R, Cc = dy.shape
    dx = torch.empty_like(dy)
    dxd = torch.empty_like(dy) if (want_drop and drop_pre_p > 0.0) else None
    capi.call("xva_layernorm_bwd", _p(dy), _p(saved["pre"]), _p(saved["mean"]), _p(saved["rstd"]), _p(gamma), _p(lens),
              Z, R, Cc, _p(dx), _p(dxd), _p(dgamma), _p(dbeta), _p(dbias), float(drop_post_p), int(seed_post),
              float(drop_pre_p), int(seed_pre), _p(seed_dev), int(relu_gate), _stream())
    return dx, (dxd if dxd is not None else dx)


def layernorm_fwd(pre, gamma, beta, lens, eps=1e-5):
    """LayerNorm of a stored pre-LN tensor [Z,R,C] -> (y, {"pre", "mean", "rstd"}); rows >= lens[z] of y are zero."""
    Z, R, Cc = pre.shape
    assert pre.is_contiguous()
    y = torch.empty_like(pre)
    mean = torch.empty(Z * R, device=pre.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    capi.call("xva_layernorm_fwd", _p(pre), _p(gamma), _p(beta), _p(lens), Z, R, Cc, float(eps), _p(y), _p(mean), _p(rstd),
              _stream())
    return y, {"pre": pre, "mean": mean, "rstd": rstd}


def colsum_(x2d_rows, C_, ld, x, out):
    capi.call("xva_colsum", _p(x), int(x2d_rows), int(C_), int(ld), _p(out), _stream())


def colsum_items_(x, out):
    """out[z, c] += sum_t x[z, t, c]; x [Z, T, C] (row / item strides free, contiguous columns), out [Z, C] view."""
    _check3(x, "x")
    Z, T, C_ = x.shape
    assert out.shape == (Z, C_) and out.stride(1) == 1
    capi.call("xva_colsum_items", _p(x), Z, T, C_, x.stride(1), x.stride(0), _p(out), out.stride(0), _stream())


def gated_act(x_in, H):
    """tanh(x_in[..., :H]) * sigmoid(x_in[..., H:2H]), tf32-rounded (python/xvapitch/wavenet.py:6-13)."""
    _check3(x_in, "x_in")
    B, T, W = x_in.shape
    assert W >= 2 * H and x_in.stride(0) == T * x_in.stride(1)
    out = torch.empty(B, T, H, device=x_in.device, dtype=torch.float32)
    capi.call("xva_gated_act_fwd", _p(x_in), B * T, H, x_in.stride(1), _p(out), _stream())
    return out


def gated_act_bwd(dacts, x_in, H):
    _check3(x_in, "x_in")
    B, T, W = x_in.shape
    assert dacts.is_contiguous() and dacts.shape == (B, T, H) and x_in.stride(0) == T * x_in.stride(1)
    out = torch.empty(B, T, 2 * H, device=x_in.device, dtype=torch.float32)
    capi.call("xva_gated_act_bwd", _p(dacts), _p(x_in), B * T, H, x_in.stride(1), _p(out), _stream())
    return out


def vits_logp(z_p, m_p, logs_p):
    """Alignment log-likelihoods of xVAPitch.train_step (xvapitch/model.py:766-771), channels-last inputs z_p [B, Ts, C],
    m_p / logs_p [B, Tt, C] -> logp [B, Ts, Tt] (frames x tokens: the layout xva_mas_width1 searches), exact fp32."""
    B, Ts, C_ = z_p.shape
    Tt = m_p.shape[1]
    assert z_p.is_contiguous() and m_p.is_contiguous() and logs_p.is_contiguous() and logs_p.shape == m_p.shape
    K = 2 * C_ + 32
    tok = torch.empty(B, Tt, K, device=z_p.device, dtype=torch.float32)
    frm = torch.empty(B, Ts, K, device=z_p.device, dtype=torch.float32)
    capi.call("xva_vits_logp_operands", _p(m_p), _p(logs_p), _p(z_p), B, Tt, Ts, C_, _p(tok), _p(frm), _stream())
    return bmm_nt(frm, tok, ref=True)


def vits_kl(z_p, logs_q, m_p, logs_p, lens, scale=1.0):
    """VitsGeneratorLoss.kl_loss (xvapitch/losses.py:86-103) on channels-last [B, T, C] tensors -> (loss as a 0-dim device
    double, (dz_p, dlogs_q, dm_p, dlogs_p) = scale * d loss / d input)."""
    B, T, C_ = z_p.shape
    for t in (z_p, logs_q, m_p, logs_p):
        assert t.is_contiguous() and t.shape == z_p.shape and t.dtype == torch.float32
    assert lens.dtype == torch.int32
    acc = torch.zeros(1, device=z_p.device, dtype=torch.float64)
    grads = tuple(torch.empty_like(z_p) for _ in range(4))
    capi.call("xva_vits_kl", _p(z_p), _p(logs_q), _p(m_p), _p(logs_p), _p(lens), B, T, C_, float(scale), _p(acc),
              *[_p(g) for g in grads], _stream())
    return acc[0] / lens.sum().to(torch.float64), grads


def vits_sample(stats, eps, lens):
    """z = (mean + eps * exp(log_scale)) * mask, stats [B, T, 2C] = [mean | log_scale] (xvapitch/model.py:1473-1474)."""
    B, T, C2 = stats.shape
    assert stats.is_contiguous() and eps.is_contiguous() and eps.shape == (B, T, C2 // 2) and lens.dtype == torch.int32
    z = torch.empty(B, T, C2 // 2, device=stats.device, dtype=torch.float32)
    capi.call("xva_vits_sample_fwd", _p(stats), _p(eps), _p(lens), B, T, C2 // 2, _p(z), _stream())
    return z


def vits_sample_bwd(dz, eps, stats, lens):
    B, T, C2 = stats.shape
    assert dz.is_contiguous() and dz.shape == (B, T, C2 // 2) and dz.dtype == torch.float32
    out = torch.empty_like(stats)
    capi.call("xva_vits_sample_bwd", _p(dz), _p(eps), _p(stats), _p(lens), B, T, C2 // 2, _p(out), _stream())
    return out


def embed_pos(tokens, emb, inp, lens, inv_freq, B, T, Cc):
    """inv_freq = None: the embedding lookup alone (no positional term)."""
    dev = inv_freq.device if inv_freq is not None else (emb if emb is not None else inp).device
    out = torch.empty(B, T, Cc, device=dev, dtype=torch.float32)
    capi.call("xva_embed_pos", _p(tokens), _p(emb), _p(inp), _p(lens), _p(inv_freq), B, T, Cc, _p(out), _stream())
    return out


def embed_bwd_(tokens, dout, demb):
    B, T, Cc = dout.shape
    capi.call("xva_embed_bwd", _p(tokens), _p(dout), B, T, Cc, _p(demb), _stream())


def scalar_conv_add_(io, x, w, bias, lens=None):
    B, T, Cc = io.shape
    capi.call("xva_scalar_conv_add", _p(io), _p(x), _p(w), _p(bias), _p(lens), B, T, Cc, _stream())


def scalar_conv_bwd_(dout, x, dw, dbias):
    B, T, Cc = dout.shape
    capi.call("xva_scalar_conv_bwd", _p(dout), _p(x), B, T, Cc, _p(dw), _p(dbias), _stream())


def rowdot_fwd(x, w, bias, lens):
    Z, R, Cc = x.shape
    out = torch.empty(Z, R, device=x.device, dtype=torch.float32)
    capi.call("xva_rowdot_fwd", _p(x), _p(w), _p(bias), _p(lens), Z, R, Cc, _p(out), _stream())
    return out


def rowdot_bwd(dout, x, w, lens, dw, db):
    Z, R, Cc = x.shape
    dx = torch.empty_like(x)
    capi.call("xva_rowdot_bwd", _p(dout), _p(x), _p(w), _p(lens), Z, R, Cc, _p(dx), _p(dw), _p(db), _stream())
    return dx


# ---------------------------------------------------------------------------------------------- losses / optimizer
def mel_mse(pred, tgt, acc):
    B, T_out, Cc = pred.shape
    capi.call("xva_mel_mse", _p(pred), _p(tgt), B, T_out, tgt.shape[2], Cc, _p(acc), _stream())


def mel_mse_grad(pred, tgt, acc, scale, ldd):
    B, T_out, Cc = pred.shape
    d = torch.empty(B, T_out, ldd, device=pred.device, dtype=torch.float32)
    capi.call("xva_mel_mse_grad", _p(pred), _p(tgt), B, T_out, tgt.shape[2], Cc, int(ldd), _p(acc), float(scale), _p(d),
              _stream())
    return d


def lens_mse(pred, tgt, lens, acc, log1p_tgt=False):
    B, T = pred.shape
    capi.call("xva_lens_mse", _p(pred), _p(tgt), _p(lens), B, T, int(log1p_tgt), _p(acc), _stream())


def lens_mse_grad(pred, tgt, lens, acc, scale, log1p_tgt=False):
    B, T = pred.shape
    d = torch.empty_like(pred)
    capi.call("xva_lens_mse_grad", _p(pred), _p(tgt), _p(lens), B, T, int(log1p_tgt), _p(acc), float(scale), _p(d),
              _stream())
    return d


def counter_add_(counter, inc=1):
    capi.call("xva_counter_add", _p(counter), int(inc), _stream())


def grad_sqnorm(g, chunks, n_chunks, out):
    capi.call("xva_grad_sqnorm", _p(g), _p(chunks), int(n_chunks), _p(out), _stream())


def lamb_step(p, g, m, v, chunks, n_chunks, norms, gnorm_sq, max_norm, lr_dev, beta1, beta2, eps, weight_decay,
              p_tf32=None):
    capi.call("xva_lamb_step", _p(p), _p(g), _p(m), _p(v), _p(chunks), int(n_chunks), _p(norms), _p(gnorm_sq),
              float(max_norm), _p(lr_dev), float(beta1), float(beta2), float(eps), float(weight_decay), _p(p_tf32),
              _stream())


def round_tf32_(src, dst):
    """dst = src rounded to tf32 (nearest); flat fp32 tensors of equal length."""
    capi.call("xva_round_tf32", _p(src), _p(dst), int(src.numel()), _stream())


# ---------------------------------------------------------------------------------------------- HiFi-GAN element-wise
def mean3_lrelu(y0, y1, y2, slope):
    out = torch.empty_like(y0)
    capi.call("xva_mean3_lrelu", _p(y0), _p(y1), _p(y2), y0.numel(), float(slope), _p(out), _stream())
    return out


def sum3(a, b, c):
    out = torch.empty_like(a)
    capi.call("xva_sum3", _p(a), _p(b), _p(c), a.numel(), _p(out), _stream())
    return out


def tanh_bwd(dy, y, ld=32):
    """dy, y [rows] -> [rows, ld] with d(pre-tanh) in column 0 and zeros elsewhere."""
    rows = y.numel()
    out = torch.empty(rows, ld, device=y.device, dtype=torch.float32)
    capi.call("xva_tanh_bwd", _p(dy), _p(y), rows, int(ld), _p(out), _stream())
    return out


def adamw_step_(p, g, m, v, lr_dev, beta1, beta2, eps, weight_decay, step, step_dev=None):
    capi.call("xva_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(lr_dev), float(beta1), float(beta2),
              float(eps), float(weight_decay), int(step), _p(step_dev), _stream())


# ---------------------------------------------------------------------------------------------- mel / losses
def reflect_pad(y, pad):
    B, n = y.shape
    out = torch.empty(B, n + 2 * pad, device=y.device, dtype=torch.float32)
    capi.call("xva_reflect_pad_fwd", _p(y), B, n, int(pad), _p(out), _stream())
    return out


def reflect_pad_bwd(dyp, n, pad):
    B = dyp.shape[0]
    dy = torch.empty(B, n, device=dyp.device, dtype=torch.float32)
    capi.call("xva_reflect_pad_bwd", _p(dyp), B, int(n), int(pad), _p(dy), _stream())
    return dy


def spec_mag(spec, nb, ld_m, eps):
    rows = spec.numel() // spec.shape[-1]
    mag = torch.empty(*spec.shape[:-1], ld_m, device=spec.device, dtype=torch.float32)
    capi.call("xva_spec_mag_fwd", _p(spec), rows, int(nb), spec.shape[-1], int(ld_m), float(eps), _p(mag), _stream())
    return mag


def spec_mag_bwd(dmag, spec, nb, eps):
    rows = spec.numel() // spec.shape[-1]
    dspec = torch.empty_like(spec)
    capi.call("xva_spec_mag_bwd", _p(dmag), _p(spec), rows, int(nb), spec.shape[-1], dmag.shape[-1], float(eps), _p(dspec),
              _stream())
    return dspec


def log_clamp(x, lo):
    out = torch.empty_like(x)
    capi.call("xva_log_clamp_fwd", _p(x), x.numel(), float(lo), _p(out), _stream())
    return out


def log_clamp_bwd(dy, x, lo):
    dx = torch.empty_like(x)
    capi.call("xva_log_clamp_bwd", _p(dy), _p(x), x.numel(), float(lo), _p(dx), _stream())
    return dx


def reduce_l1(a, b, acc):
    """acc (device double, 0-dim or 1-element view) += sum |a - b|"""
    capi.call("xva_reduce_loss", _p(a), _p(b), a.numel(), 0, 0.0, _p(acc), _stream())


def reduce_sq(a, c, acc):
    """acc += sum (c - a)^2"""
    capi.call("xva_reduce_loss", _p(a), None, a.numel(), 1, float(c), _p(acc), _stream())


def l1_grad(a, b, scale, out=None, gate_slope=1.0):
    """scale * d(sum |a - b|)/db = scale * sign(b - a), times gate_slope where b <= 0; accumulated into ``out`` when
    given."""
    acc = out is not None
    if out is None:
        out = torch.empty_like(b)
    capi.call("xva_loss_grad", _p(a), _p(b), b.numel(), 0, 0.0, float(scale), float(gate_slope), int(acc), _p(out),
              _stream())
    return out


def l1_loss_grad(a, b, scale, acc, gate_slope=1.0):
    """acc += sum |a - b| and returns scale * sign(b - a) (times gate_slope where b <= 0) in one pass."""
    out = torch.empty_like(b)
    capi.call("xva_l1_loss_grad", _p(a), _p(b), b.numel(), float(scale), float(gate_slope), _p(acc), _p(out), _stream())
    return out


def sq_grad(a, c, scale, out=None, accumulate=True):
    """scale * d(sum (c - a)^2)/da = 2 scale (a - c); accumulated into ``out`` when given (unless accumulate=False)."""
    acc = out is not None and accumulate
    if out is None:
        out = torch.empty_like(a)
    capi.call("xva_loss_grad", _p(a), None, a.numel(), 1, float(c), float(scale), 1.0, int(acc), _p(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------- discriminator pieces
def conv_c1_fwd(wave, geom, w, bias, k, s, pad, Z, Lout, Lout_p, Cout, slope):
    """geom = (xs_b, xs_q, xs_c, P, Lsrc, L): how the Z = B*P sequences are cut out of wave [B, Lsrc] (include/xva_b200.h)."""
    out = torch.empty(Z, Lout_p, Cout, device=wave.device, dtype=torch.float32)
    xs_b, xs_q, xs_c, P, Lsrc, L = geom
    capi.call("xva_conv_c1_fwd", _p(wave), xs_b, xs_q, xs_c, P, Lsrc, L, _p(w), _p(bias), k, s, pad, Z, Lout, Lout_p, Cout,
              float(slope), _p(out), _stream())
    return out


def conv_c1_bwd_w(dpre, wave, geom, k, s, pad, Lout, dw, db):
    Z, Lout_p, Cout = dpre.shape
    xs_b, xs_q, xs_c, P, Lsrc, L = geom
    capi.call("xva_conv_c1_bwd_w", _p(dpre), _p(wave), xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout,
              _p(dw), _p(db), _stream())


def conv_c1_bwd_x(dpre, w, geom, k, s, pad, Lout, scale, dwave):
    Z, Lout_p, Cout = dpre.shape
    xs_b, xs_q, xs_c, P, Lsrc, L = geom
    capi.call("xva_conv_c1_bwd_x", _p(dpre), _p(w), xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout,
              float(scale), _p(dwave), _stream())


def avgpool4(x):
    B, L = x.shape
    out = torch.empty(B, L // 2 + 1, device=x.device, dtype=torch.float32)
    capi.call("xva_avgpool4_fwd", _p(x), B, L, _p(out), _stream())
    return out


def avgpool4_bwd(dout, L):
    B = dout.shape[0]
    dx = torch.empty(B, L, device=dout.device, dtype=torch.float32)
    capi.call("xva_avgpool4_bwd", _p(dout), B, L, _p(dx), _stream())
    return dx


def zero_tail_rows_(x, Lvalid):
    Z, Lp, Cc = x.shape
    capi.call("xva_zero_tail_rows", _p(x), Z, Lp, int(Lvalid), Cc, _stream())


# ---------------------------------------------------------------------------------------------- xVAPitch text encoder
def text_embed(tokens, emb, lang, lens, scale, ld, want_x_emb=True):
    """TextEncoder.forward's input stage (python/xvapitch/model.py:1152-1165): tokens int64 [B, T], emb [V, C], lang [B, L]
    or None, lens int32 [B] -> (x [B, T, ld] = [emb[tokens] * scale | lang | 0] on rows t < lens[b], zero rows after;
    x_emb [B, T, C] = emb[tokens] * scale at every position, or None)."""
    B, T = tokens.shape
    V, C_ = emb.shape
    L = 0 if lang is None else lang.shape[1]
    assert tokens.dtype == torch.int64 and tokens.is_contiguous() and emb.is_contiguous() and lens.dtype == torch.int32
    assert lang is None or (lang.is_contiguous() and lang.shape[0] == B)
    out = torch.empty(B, T, ld, device=emb.device, dtype=torch.float32)
    x_emb = torch.empty(B, T, C_, device=emb.device, dtype=torch.float32) if want_x_emb else None
    capi.call("xva_text_embed_fwd", _p(tokens), _p(emb), _p(lang), _p(lens), B, T, C_, L, int(ld), float(scale), _p(out),
              _p(x_emb), _stream())
    return out, x_emb


def text_embed_bwd_(tokens, dout, lens, C_, scale, demb):
    """demb[tokens[b, t], :C] += scale * dout[b, t, :C] on rows t < lens[b] (lens None: every row). dout [B, T, >= C]."""
    _check3(dout, "dout")
    B, T, _ = dout.shape
    assert dout.stride(0) == T * dout.stride(1) and demb.is_contiguous() and demb.shape[1] == C_
    capi.call("xva_text_embed_bwd", _p(tokens), _p(dout), _p(lens), B, T, int(C_), dout.stride(1), float(scale), _p(demb),
              _stream())


def rel_band_add_(s, rel, T, W):
    """s[z, t, t + r - W] += rel[z, t, r], r in [0, 2W], inside [0, T) (glow_tts.py:178-186). s [Z, T, ld], rel [Z, T, ldr]."""
    Z, R, ld = s.shape
    assert s.is_contiguous() and rel.is_contiguous() and rel.shape[:2] == (Z, R) and R == T
    capi.call("xva_rel_band_add", _p(s), _p(rel), Z, int(T), int(W), ld, rel.shape[2], _stream())
    return s


def rel_band_gather(p, T, W, ldo=32):
    """out[z, t, r] = p[z, t, t + r - W] inside the band and [0, T), zero elsewhere (glow_tts.py:192-195); p [Z, T, ld] ->
    [Z, T, ldo], tf32-rounded."""
    Z, R, ld = p.shape
    assert p.is_contiguous() and R == T and ldo >= 2 * W + 1
    out = torch.empty(Z, T, ldo, device=p.device, dtype=torch.float32)
    capi.call("xva_rel_band_gather", _p(p), Z, int(T), int(W), ld, int(ldo), _p(out), _stream())
    return out


def pad_cols(x, ld):
    """[.., C] contiguous -> [.., ld] with zero pad columns (a row pitch the MN-major GEMM operands accept)."""
    assert x.is_contiguous() and ld >= x.shape[-1]
    C_ = x.shape[-1]
    out = torch.empty(*x.shape[:-1], ld, device=x.device, dtype=torch.float32)
    capi.call("xva_pad_cols", _p(x), x.numel() // C_, C_, int(ld), _p(out), _stream())
    return out
