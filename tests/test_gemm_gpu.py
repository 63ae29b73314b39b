"""Parity of the tcgen05 tap-GEMM (xva_gemm) against (a) plain torch fp32 math with TF32 off and (b) the exact-fp32
SIMT checker xva_gemm_ref, for every mode / epilogue the FastPitch and HiFi-GAN layers use.

Tolerance: operands are read as tf32 (10-bit mantissa) with fp32 accumulation, so the bar for the tensor-core
kernel is relative L2 error <= 2e-3 per tensor; the SIMT checker must agree with torch to 1e-5.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_TC = 2e-3
TOL_REF = 2e-5


def rel(a, b):
    r = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    if not (r < TOL_TC):
        _describe(a, b)
    return r


def _describe(got, want):
    """Error anatomy printed on failure: which rows / columns are off tells descriptor bugs from wiring bugs."""
    try:
        g, w = got.reshape(-1, got.shape[-1]).float(), want.reshape(-1, want.shape[-1]).float()
        print(f"\n[describe] shape {tuple(got.shape)} max|got| {g.abs().max():.4g} max|want| {w.abs().max():.4g} "
              f"nan {int(torch.isnan(g).sum())} zeros {(g == 0).float().mean():.3f}")
        err = (g - w).abs()
        rows = err.mean(1)
        cols = err.mean(0)
        print("[describe] row-mean err, first 16 rows:", [f"{v:.2g}" for v in rows[:16].tolist()])
        print("[describe] row-mean err by row%8      :", [f"{rows[i::8].mean():.2g}" for i in range(8)])
        nb = (cols.numel() + 31) // 32
        print("[describe] col-mean err per 32-col blk:", [f"{cols[i * 32:(i + 1) * 32].mean():.2g}" for i in range(min(nb, 16))])
        print("[describe] col-mean err by col%8 (blk0):", [f"{cols[i:32:8].mean():.2g}" for i in range(8)])
        rb = (rows.numel() + 127) // 128
        print("[describe] row-mean err per 128-row blk:", [f"{rows[i * 128:(i + 1) * 128].mean():.2g}" for i in range(min(rb, 12))])
    except Exception as e:  # pragma: no cover
        print("[describe] failed:", e)


@pytest.fixture(scope="module", autouse=True)
def _strict_fp32(lib):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _ops():
    from xva_trainer_b200 import ops

    return ops


def conv_ref(x, wp, shifts):
    """out[b,t,n] = sum_j x[b,t+s_j,:] @ wp[j].T with zero rows outside [0,T)."""
    B, T, K = x.shape
    out = torch.zeros(B, T, wp.shape[1], device=x.device, dtype=torch.float64)
    xd, wd = x.double(), wp.double()
    for j, s in enumerate(shifts):
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            out[:, lo:hi] += xd[:, lo + s:hi + s] @ wd[j].T
    return out.float()


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("B,T,K,N,shifts", [
    (2, 200, 64, 64, (0,)),
    (3, 130, 384, 192, (0,)),          # qkv projection shape
    (2, 257, 384, 80, (0,)),           # proj 384 -> 80
    (2, 300, 384, 1536, (-1, 0, 1)),   # ConvFF first conv
    (2, 140, 1536, 384, (-1, 0, 1)),   # ConvFF second conv
    (2, 96, 80, 512, (-3, -2, -1, 0, 1, 2, 3)),  # conv_pre k7, K not a multiple of 32
    (1, 500, 128, 128, (-5, 0, 5)),    # dilated k3 d5
    (2, 64, 32, 32, (-1, 0, 1)),
])
def test_conv_fwd(B, T, K, N, shifts):
    ops = _ops()
    x, w = gen(B, T, K, seed=1), gen(len(shifts), N, K, seed=2, scale=K ** -0.5)
    bias = gen(N, seed=3)
    want = conv_ref(x, w, shifts) + bias
    got_ref = ops.conv_fwd(x, w, shifts, bias=bias, ref=True)
    got = ops.conv_fwd(x, w, shifts, bias=bias)
    torch.cuda.synchronize()
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC, f"tcgen05 rel err {rel(got, want)}"


def test_conv_fwd_relu_lens_strided():
    ops = _ops()
    B, T, K, N = 3, 200, 384, 256
    big = gen(B, T, K + 64, seed=4)
    x = big[:, :, 32:32 + K]  # strided view: row stride K+64
    w = gen(3, N, K, seed=5, scale=K ** -0.5)
    lens = torch.tensor([200, 77, 128], device="cuda", dtype=torch.int32)
    want = torch.relu(conv_ref(x.contiguous(), w, (-1, 0, 1)))
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).float()[..., None]
    want = want * mask
    got = ops.conv_fwd(x, w, (-1, 0, 1), relu=True, lens=lens)
    got_ref = ops.conv_fwd(x, w, (-1, 0, 1), relu=True, lens=lens, ref=True)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC


@pytest.mark.parametrize("N,K,shifts", [(384, 64, (0,)), (384, 1536, (-1, 0, 1)), (256, 384, (-1, 0, 1))])
def test_conv_fwd_layernorm(N, K, shifts):
    ops = _ops()
    B, T = 2, 150
    x = gen(B, T, K, seed=6)
    w = gen(len(shifts), N, K, seed=7, scale=K ** -0.5)
    bias, res = gen(N, seed=8), gen(B, T, N, seed=9)
    gamma, beta = 1 + 0.1 * gen(N, seed=10), 0.1 * gen(N, seed=11)
    lens = torch.tensor([150, 99], device="cuda", dtype=torch.int32)
    pre = conv_ref(x, w, shifts) + bias + res
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).float()[..., None]
    want = torch.nn.functional.layer_norm(pre, (N,), gamma, beta, 1e-5) * mask
    for ref in (True, False):
        got, extra = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, ln=(gamma, beta), lens=lens, save_ln=True,
                                  ref=ref)
        tol = TOL_REF if ref else TOL_TC
        assert rel(got, want) < tol, f"ref={ref} out {rel(got, want)}"
        assert rel(extra["pre"], pre) < tol
        assert rel(extra["mean"].view(B, T), pre.mean(-1)) < 5e-3 + tol
        assert rel(extra["rstd"].view(B, T), (pre.var(-1, unbiased=False) + 1e-5).rsqrt()) < tol


def test_relu_then_layernorm_predictor():
    # ConvReLUNorm: LN(relu(conv(x)+b))  (common/layers.py:94-97)
    ops = _ops()
    B, T, K, N = 2, 160, 384, 256
    x, w, bias = gen(B, T, K, seed=12), gen(3, N, K, seed=13, scale=K ** -0.5), gen(N, seed=14)
    gamma, beta = 1 + 0.1 * gen(N, seed=15), 0.1 * gen(N, seed=16)
    want = torch.nn.functional.layer_norm(torch.relu(conv_ref(x, w, (-1, 0, 1)) + bias), (N,), gamma, beta, 1e-5)
    got = ops.conv_fwd(x, w, (-1, 0, 1), bias=bias, relu=True, ln=(gamma, beta))
    assert rel(got, want) < TOL_TC


@pytest.mark.parametrize("B,T,K,N,shifts", [
    (2, 200, 384, 1536, (-1, 0, 1)),
    (2, 300, 64, 384, (0,)),
    (3, 130, 384, 192, (0,)),
    (1, 400, 128, 128, (-3, 0, 3)),
])
def test_conv_dgrad(B, T, K, N, shifts):
    ops = _ops()
    dy, w = gen(B, T, N, seed=20), gen(len(shifts), N, K, seed=21, scale=N ** -0.5)
    x = gen(B, T, K, seed=22).requires_grad_(True)
    (conv_ref_autograd(x, w, shifts) * dy).sum().backward()
    want = x.grad
    got_ref = ops.conv_dgrad(dy, w, shifts, ref=True)
    got = ops.conv_dgrad(dy, w, shifts)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC, f"dgrad rel err {rel(got, want)}"


def conv_ref_autograd(x, wp, shifts):
    B, T, K = x.shape
    out = 0
    for j, s in enumerate(shifts):
        xs = torch.zeros_like(x)
        lo, hi = max(0, -s), min(T, T - s)
        pad = torch.nn.functional.pad(x, (0, 0, max(0, -s), max(0, s)))
        xs = pad[:, max(0, s):max(0, s) + T]
        out = out + xs @ wp[j].T
    return out


def test_dgrad_relu_gate_residual():
    ops = _ops()
    B, T, K, N = 2, 130, 1536, 384
    dy, w = gen(B, T, N, seed=23), gen(3, N, K, seed=24, scale=N ** -0.5)
    h = gen(B, T, K, seed=25)
    want = conv_ref(dy, w.transpose(1, 2).contiguous(), (1, 0, -1)) * (h > 0).float()
    got = ops.conv_dgrad(dy, w, (-1, 0, 1), gate=h)
    got_ref = ops.conv_dgrad(dy, w, (-1, 0, 1), gate=h, ref=True)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC


@pytest.mark.parametrize("B,T,K,N,shifts,split", [
    (4, 200, 384, 1536, (-1, 0, 1), 2),
    (2, 300, 64, 384, (0,), 1),
    (3, 130, 384, 192, (0,), 3),
    (2, 90, 1536, 384, (-1, 0, 1), None),
])
def test_conv_wgrad(B, T, K, N, shifts, split):
    ops = _ops()
    dy, x = gen(B, T, N, seed=30), gen(B, T, K, seed=31)
    w = gen(len(shifts), N, K, seed=32).requires_grad_(True)
    (conv_ref_autograd(x, w, shifts) * dy).sum().backward()
    want = w.grad
    got_ref = ops.conv_wgrad(dy, x, shifts, ref=True)
    got = ops.conv_wgrad(dy, x, shifts, split=split)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC, f"wgrad rel err {rel(got, want)}"


def test_attention_bmms_on_qkv_slices():
    ops = _ops()
    B, T, D = 3, 200, 64
    qkv = gen(B, T, 3 * D, seed=40)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    s_want = torch.bmm(q, k.transpose(1, 2)) * 0.125
    s_got = ops.bmm_nt(q, k, alpha=0.125)
    assert rel(s_got, s_want) < TOL_TC
    # key dimension T = 200 is not a multiple of 32: the probabilities live in rows of stride 224 (zero pad columns),
    # which is how the engine lays out the score tensor (include/xva_b200.h, MN-major operand rule)
    p_pad = torch.zeros(B, T, 224, device="cuda")
    p = p_pad[..., :T]
    p.copy_(torch.softmax(s_want, -1))
    o_want = torch.bmm(p, v)
    o_got = ops.bmm_nn(p, v)
    assert rel(o_got, o_want) < TOL_TC
    do = gen(B, T, D, seed=41)
    dv_want = torch.bmm(p.transpose(1, 2), do)
    dv_got = ops.bmm_tn(p, do)
    assert rel(dv_got, dv_want) < TOL_TC
    for fn, want in ((lambda r: ops.bmm_nt(q, k, alpha=0.125, ref=r), s_want), (lambda r: ops.bmm_nn(p, v, ref=r), o_want),
                     (lambda r: ops.bmm_tn(p, do, ref=r), dv_want)):
        assert rel(fn(True), want) < TOL_REF


def test_dropout_matches_checker_and_keeps_scale():
    ops = _ops()
    B, T, K, N = 2, 256, 384, 384
    x, w = gen(B, T, K, seed=50), gen(1, N, K, seed=51, scale=K ** -0.5)
    a = ops.conv_fwd(x, w, (0,), drop_p=0.1, seed=1234)
    b = ops.conv_fwd(x, w, (0,), drop_p=0.1, seed=1234, ref=True)
    full = ops.conv_fwd(x, w, (0,), ref=True)
    assert torch.equal(a == 0, b == 0)           # same counter-based mask in both implementations
    frac = (a == 0).float().mean().item()
    assert 0.08 < frac < 0.12
    kept = a != 0
    assert rel(a[kept], full[kept] / 0.9) < TOL_TC
    c = ops.conv_fwd(x, w, (0,), drop_p=0.1, seed=99)
    assert not torch.equal(a == 0, c == 0)


def test_full_size_linearity_property():
    """BASELINE size (B=32, T=880): f(x1 + x2) == f(x1) + f(x2) for the k3 conv, checked without any reference."""
    ops = _ops()
    B, T, K, N = 32, 880, 384, 1536
    x1, x2 = gen(B, T, K, seed=60), gen(B, T, K, seed=61)
    w = gen(3, N, K, seed=62, scale=K ** -0.5)
    y12 = ops.conv_fwd(x1 + x2, w, (-1, 0, 1))
    y1, y2 = ops.conv_fwd(x1, w, (-1, 0, 1)), ops.conv_fwd(x2, w, (-1, 0, 1))
    assert rel(y12, y1 + y2) < TOL_TC
    # and a sampled exact check of 64 rows against fp64 math
    rows = torch.randint(0, T, (64,), device="cuda")
    want = conv_ref(x1[:2], w, (-1, 0, 1))[:, rows]
    assert rel(y1[:2][:, rows], want) < TOL_TC


# ------------------------------------------------------------------------------------------------ segmented row tiles
@pytest.mark.parametrize("B,T,K,N,shifts", [
    (32, 160, 384, 192, (0,)),            # FastPitch encoder: 160 tokens = 5 segments of 32, tiles span items
    (9, 160, 96, 256, (-1, 0, 1)),        # odd item count: last tile half empty, CTA-pair duplicate
    (22, 10, 128, 128, (-2, -1, 0, 1, 2)),  # DiscriminatorP period 11: 10 rows per sequence
    (7, 51, 64, 96, (-1, 0, 1)),          # 51 rows -> 64-row segments
    (5, 83, 64, 64, (-1, 0, 1)),          # 83 rows -> three 32-row segments
    (8, 48, 64, 96, (-1, 0, 1)),          # 48 rows -> two 32-row segments, the second half empty
    (8, 880, 32, 64, (-2, 0, 2)),         # FastPitch decoder rows (880 = 6 x 128 + 112)
])
def test_segmented_tiles(B, T, K, N, shifts):
    """Short sequences share 128-row tiles (segments of 32 / 64 rows): conv halo stays per item (TMA zero fill per
    segment), every epilogue input (bias, gate, residual, lens) follows the row's own item."""
    ops = _ops()
    x, w = gen(B, T, K, seed=41), gen(len(shifts), N, K, seed=42, scale=K ** -0.5)
    bias, res, gate = gen(N, seed=43), gen(B, T, N, seed=44), gen(B, T, N, seed=45)
    lens = torch.randint(1, T + 1, (B,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(46)).int()
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).float()[..., None]
    want = ((conv_ref(x, w, shifts) + bias) * torch.where(gate > 0, 1.0, 0.1) + res) * mask
    got_ref = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, gate=gate, gate_slope=0.1, lens=lens, ref=True)
    got = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, gate=gate, gate_slope=0.1, lens=lens)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC
    # dgrad through the same tiles (MN-major B)
    dy = gen(B, T, N, seed=47)
    want_dx = conv_ref(dy, w.transpose(1, 2).contiguous(), [-s for s in shifts])
    got_dx = ops.conv_dgrad(dy, w, shifts)
    assert rel(got_dx, want_dx) < TOL_TC
    # LayerNorm epilogue
    if N % 16 == 0:
        gamma, beta = 1 + 0.1 * gen(N, seed=48), 0.1 * gen(N, seed=49)
        pre = conv_ref(x, w, shifts) + bias + res
        want_ln = torch.nn.functional.layer_norm(pre, (N,), gamma, beta, 1e-5) * mask
        got_ln, sv = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, ln=(gamma, beta), save_ln=True, lens=lens)
        assert rel(got_ln, want_ln) < TOL_TC
        assert rel(sv["pre"], pre) < TOL_TC
        assert rel(sv["mean"].view(B, T), pre.mean(-1)) < 2e-3 + 1e-3


# ------------------------------------------------------------------------------------------------ grouped convolutions
def _grouped_case(Cin, Cout, G, k, stride, seed):
    """DiscriminatorS-style grouped strided conv: returns everything the three launches need plus the torch result."""
    import math
    B, L = 3, 4 * 97
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, L, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin // G, k, device="cuda", generator=g) * (Cin // G * k) ** -0.5
    pad = (k - 1) // 2
    y = torch.nn.functional.conv1d(x.transpose(1, 2), w, None, stride=stride, padding=pad, groups=G).transpose(1, 2).contiguous()
    return B, L, x, w, pad, y


@pytest.mark.parametrize("Cin,Cout,G,k,stride", [(128, 128, 4, 41, 2), (128, 256, 16, 41, 2), (256, 512, 16, 41, 4),
                                                   (64, 128, 2, 5, 1)])
def test_grouped_conv_single_launch(Cin, Cout, G, k, stride):
    """Grouped conv forward / input gradient / weight gradient, one launch each (xva_gemm_args.groups), through the
    packing hifigan._Disc uses (phase-major taps on the [L/stride, stride*Cin] view, groups with < 32 channels merged
    into block-diagonal super-groups) vs torch's grouped conv1d and its autograd."""
    from xva_trainer_b200 import hifigan as hg
    ops = _ops()
    B, L, x, w, pad, y = _grouped_case(Cin, Cout, G, k, stride, seed=50 + G)
    m = hg._DiscConv(Cin, Cout, k, stride, pad, groups=G).cuda()
    Gp, Ogp, Cgp, f = hg._Disc._group_geom(None, m)
    taps = m.taps()
    order = torch.tensor([j for j, _, _ in taps], device="cuda")
    Og, Cg = Cout // G, Cin // G
    wk = w.index_select(2, order).permute(2, 0, 1)
    if f > 1:
        slot = (torch.arange(Cout, device="cuda") // Og) % f
        mask = (slot[:, None] == torch.arange(f, device="cuda")[None, :]).float()[:, :, None]
        wk = (wk.unsqueeze(2) * mask).reshape(k, Cout, Cgp)
    wk = wk.contiguous()
    xv = x.view(B, L // stride, stride * Cin)
    shifts = [sh for _, sh, _ in taps]
    a_cols = [ph * Cin for _, _, ph in taps]
    Lout = y.shape[1]
    kw = dict(a_cols=a_cols, out_rows=Lout, groups=Gp, grp_step=Cgp)
    got_ref = ops.conv_fwd(xv, wk, shifts, ref=True, **kw)
    got = ops.conv_fwd(xv, wk, shifts, **kw)
    assert rel(got_ref, y) < TOL_REF
    assert rel(got, y) < TOL_TC
    # gradients
    dy = gen(B, Lout, Cout, seed=60)
    xl, wl = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yy = torch.nn.functional.conv1d(xl.transpose(1, 2), wl, None, stride=stride, padding=pad, groups=G).transpose(1, 2)
    (yy * dy).sum().backward()
    for ref in (True, False):
        tol = TOL_REF if ref else TOL_TC
        dx = torch.zeros_like(x)
        dxv = dx.view(B, L // stride, stride * Cin)
        for ph in range(stride):
            idx = [i for i, (_, _, p_) in enumerate(taps) if p_ == ph]
            if not idx:
                continue
            ops.conv_dgrad(dy, wk[idx[0]:idx[-1] + 1], [taps[i][1] for i in idx], out=dxv[..., ph * Cin:(ph + 1) * Cin],
                           out_rows=L // stride, groups=Gp, ref=ref)
        assert rel(dx, xl.grad) < tol, ("dgrad", ref)
        dw = ops.conv_wgrad(dy, xv, shifts, x_cols=a_cols, n_cols=Cgp, groups=Gp, grp_step=Cgp, ref=ref)   # [k, Cout, Cgp]
        # un-pack: diagonal blocks of the super-groups, taps back in kernel order
        if f > 1:
            dw = (dw.view(k, Cout, f, Cg) * mask).sum(2)
        dw_ref_layout = torch.empty_like(w)
        dw_ref_layout[:, :, order] = dw.permute(1, 2, 0)
        assert rel(dw_ref_layout, wl.grad) < tol, ("wgrad", ref)


@pytest.mark.parametrize("B,T,N,K,shifts,split", [(2, 4096, 32, 32, (-5, 0, 5), 24), (3, 1000, 64, 96, (-1, 0, 1), 9),
                                                  (1, 2048, 128, 128, (0,), 8)])
def test_conv_wgrad_row_chunked_split(B, T, N, K, shifts, split):
    """More CTAs per output tile than items: every item's contraction rows are cut into chunks (few long sequences, the
    HiFi-GAN generator's case); the tap shift still reads across chunk boundaries and zero-fills only at the item's ends."""
    ops = _ops()
    dy, x = gen(B, T, N, seed=71), gen(B, T, K, seed=72)
    want = torch.zeros(len(shifts), N, K, device="cuda", dtype=torch.float64)
    for j, s in enumerate(shifts):
        lo, hi = max(0, -s), min(T, T - s)
        want[j] = torch.einsum("btn,btk->nk", dy[:, lo:hi].double(), x[:, lo + s:hi + s].double())
    got = ops.conv_wgrad(dy, x, shifts, split=split)
    assert rel(got, want.float()) < TOL_TC


@pytest.mark.parametrize("B,T,K,N,shifts", [
    (2, 256, 96, 128, (-1, 0, 1)),                                   # rows a multiple of 128: un-segmented tiles
    (1, 1000, 64, 64, tuple(5 * (j - 5) for j in range(11))),        # HiFi-GAN ResBlock k = 11, dilation 5 (halo 25 rows)
    (1, 700, 128, 32, tuple(3 * (j - 3) for j in range(7))),         # k = 7, dilation 3
    (3, 384, 1536, 384, (-1, 0, 1)),                                 # ConvFF second conv: N = 384 single tile
    (16, 1280, 192, 256, (-1, 0, 1)),                                # a full wave of row tiles: CTA pairs
    (1, 300, 80, 512, (-3, -2, -1, 0, 1, 2, 3)),                     # conv_pre k = 7, K not a multiple of 32
    (2, 128, 64, 64, (0, -1)),                                       # transposed-conv phase group (two taps)
])
def test_halo_mode_taps_from_one_activation_tile(B, T, K, N, shifts):
    """(opt-in path, XVA_GEMM_HALO) k-tap convolutions on un-segmented tiles load the activation tile once per k-block with halo rows and read every tap
    through a row-shifted descriptor: forward, input gradient (negated shifts, MN-major weights), gated / residual / LN
    epilogues, per-item zero padding at both ends."""
    ops = _ops()
    x, w = gen(B, T, K, seed=81), gen(len(shifts), N, K, seed=82, scale=(K * len(shifts)) ** -0.5)
    bias, res, gate = gen(N, seed=83), gen(B, T, N, seed=84), gen(B, T, N, seed=85)
    want = (conv_ref(x, w, shifts) + bias) * torch.where(gate > 0, 1.0, 0.1) + res
    got = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, gate=gate, gate_slope=0.1, halo=True)
    got_ref = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, gate=gate, gate_slope=0.1, ref=True)
    assert rel(got_ref, want) < TOL_REF
    assert rel(got, want) < TOL_TC
    if K % 32 == 0:      # the input gradient reads the weights MN-major: rows of 32-column chunks
        dy = gen(B, T, N, seed=86)
        want_dx = conv_ref(dy, w.transpose(1, 2).contiguous(), [-s for s in shifts])
        assert rel(ops.conv_dgrad(dy, w, shifts, halo=True), want_dx) < TOL_TC
    if N % 16 == 0 and N <= 512:
        gamma, beta = 1 + 0.1 * gen(N, seed=87), 0.1 * gen(N, seed=88)
        pre = conv_ref(x, w, shifts) + bias + res
        got_ln = ops.conv_fwd(x, w, shifts, bias=bias, residual=res, ln=(gamma, beta), halo=True)
        assert rel(got_ln, torch.nn.functional.layer_norm(pre, (N,), gamma, beta, 1e-5)) < TOL_TC


@pytest.mark.parametrize("B,T,N,K,shifts", [
    (2, 4096, 32, 32, tuple(j - 5 for j in range(11))),            # ResBlock k = 11 at 32 channels: all 11 taps in one tile
    (3, 1500, 64, 64, tuple(3 * (j - 3) for j in range(7))),         # 64 channels, k = 7 dilation 3: taps in two groups
    (2, 900, 128, 128, tuple(5 * (j - 5) for j in range(11))),       # 128 channels, k = 11 dilation 5: three groups
    (4, 700, 1, 32, (-3, -2, -1, 0, 1, 2, 3)),                       # conv_post: one output channel
])
def test_conv_wgrad_multi_tap_tiles(B, T, N, K, shifts):
    """Small-channel weight gradients accumulate several taps side by side in TMEM (dy fetched once per k-block for all of
    them) and split the rows of every item over the SMs; vs the fp64 einsum."""
    ops = _ops()
    ld = 32 if N < 32 else N
    dyp = torch.zeros(B, T, ld, device="cuda")
    dyp[..., :N] = gen(B, T, N, seed=91)
    dy, x = dyp[..., :N], gen(B, T, K, seed=92)
    want = torch.zeros(len(shifts), N, K, device="cuda", dtype=torch.float64)
    for j, s in enumerate(shifts):
        lo, hi = max(0, -s), min(T, T - s)
        want[j] = torch.einsum("btn,btk->nk", dy[:, lo:hi].double(), x[:, lo + s:hi + s].double())
    got = ops.conv_wgrad(dy, x, shifts)
    got_ref = ops.conv_wgrad(dy, x, shifts, ref=True)
    assert rel(got_ref, want.float()) < TOL_REF
    assert rel(got, want.float()) < TOL_TC
