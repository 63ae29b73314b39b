"""Row kernels (csrc/rowops.cu) against plain torch fp32: masked softmax forward / backward (transformer.py:120-127),
LayerNorm forward / backward (nn.LayerNorm autograd), column sums. Outputs that feed a GEMM are stored tf32-rounded, so
the bar is relative L2 <= 5e-4 (one rounding, 2^-11); statistics and un-rounded gradients <= 1e-5. The dropout mask is a
pure function of (seed, element): forward and backward kernels must agree on it element by element."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def gen(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g)


@pytest.mark.parametrize("Z,R,N", [(3, 70, 880), (4, 33, 160), (2, 20, 100), (2, 9, 37)])
def test_softmax_fwd_bwd(lib, Z, R, N):
    from xva_trainer_b200 import ops
    ld = (N + 31) // 32 * 32
    s = torch.full((Z, R, ld), float("nan"), device="cuda")      # pad columns are never written by the GEMM
    s[..., :N] = gen(Z, R, N, seed=1) * 3
    lens = torch.randint(1, N + 1, (Z,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)).int()
    mask = torch.arange(N, device="cuda")[None, None, :] >= lens[:, None, None]
    want = torch.softmax(s[..., :N].masked_fill(mask, float("-inf")), -1)
    p, pd = ops.softmax_fwd(s, lens, N)
    assert pd is p
    assert rel(p[..., :N], want) < 5e-4
    assert float(p[..., N:].abs().max() if ld > N else 0.0) == 0.0
    # dropout: pd = p * m with m in {0, 1/(1-q)}, keep rate ~ 1-q, and backward re-derives the same mask
    q, seed = 0.25, 1234567
    p2, pd2 = ops.softmax_fwd(s, lens, N, q, seed)
    assert torch.equal(p2, p)
    live = p[..., :N] > 1e-20
    m = torch.where(live, pd2[..., :N] / p[..., :N].clamp_min(1e-30), torch.zeros_like(p[..., :N]))
    kept = m[live] > 0
    assert abs(float(kept.float().mean()) - (1 - q)) < 4 * (q * (1 - q) / kept.numel()) ** 0.5 + 2e-3
    assert float((m[live][kept] - 1 / (1 - q)).abs().max()) < 2e-3
    g = gen(Z, R, ld, seed=3)
    mfull = torch.zeros(Z, R, ld, device="cuda")
    mfull[..., :N] = torch.where(m > 0, torch.full_like(m, 1 / (1 - q)), torch.zeros_like(m))
    mfull[..., :N][~live] = 1 / (1 - q)        # entries with p == 0 contribute nothing either way
    gp = g[..., :N] * mfull[..., :N]
    dot = (want * gp).sum(-1, keepdim=True)
    want_ds = 0.125 * want * (gp - dot)
    dpd = g.clone()
    ops.softmax_bwd_(p, dpd, N, 0.125, q, seed)
    # rows where a dropped-vs-kept decision of a p == 0 entry differs do not matter: p multiplies everything
    assert rel(dpd[..., :N], want_ds) < 2e-3
    assert float(dpd[..., N:].abs().max() if ld > N else 0.0) == 0.0


@pytest.mark.parametrize("Z,R,C", [(3, 50, 384), (2, 33, 256), (2, 17, 80), (3, 41, 780), (2, 19, 708), (2, 9, 1024), (2, 21, 516)])
def test_layernorm_fwd_bwd(lib, Z, R, C):
    from xva_trainer_b200 import ops
    x = gen(Z, R, C, seed=5) * 2 + 0.5
    gamma, beta = 1 + 0.1 * gen(C, seed=6), 0.1 * gen(C, seed=7)
    lens = torch.tensor([R, max(1, R // 2), 1][:Z], device="cuda", dtype=torch.int32)
    mask = (torch.arange(R, device="cuda")[None, :] < lens[:, None]).float()[..., None]
    xl = x.clone().requires_grad_(True)
    gl, bl = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    want = torch.nn.functional.layer_norm(xl, (C,), gl, bl, 1e-5) * mask
    y, sv = ops.layernorm_fwd(x, gamma, beta, lens)
    assert rel(y, want) < 5e-4
    assert rel(sv["mean"].view(Z, R), x.mean(-1)) < 1e-5
    assert rel(sv["rstd"].view(Z, R), (x.var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-5
    dy = gen(Z, R, C, seed=8)
    (want * dy).sum().backward()
    dgamma, dbeta, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    dx, dxd = ops.layernorm_bwd(dy, sv, gamma, lens, dgamma, dbeta, dbias=dbias)
    assert dxd is dx
    assert rel(dx, xl.grad) < 5e-4
    assert rel(dgamma, gl.grad) < 1e-4 and rel(dbeta, bl.grad) < 1e-4
    assert rel(dbias, dx.sum((0, 1))) < 1e-5
    # with a pre-LN dropout site: dx_drop = dx * mask(seed), the same mask the GEMM epilogue applied in the forward
    dx2, dxd2 = ops.layernorm_bwd(dy, sv, gamma, lens, dgamma, dbeta, want_drop=True, drop_pre_p=0.3, seed_pre=99)
    ratio = dxd2[dx2.abs() > 1e-6] / dx2[dx2.abs() > 1e-6]
    kept = ratio.abs() > 0
    assert abs(float(kept.float().mean()) - 0.7) < 0.02
    assert float((ratio[kept] - 1 / 0.7).abs().max()) < 2e-3


def test_dropout_mask_shared_by_gemm_epilogue_and_layernorm_bwd(lib):
    """FFT block: the GEMM epilogue drops the branch before the residual add (transformer.py:51,139) and
    xva_layernorm_bwd must regenerate exactly that mask for the branch gradient."""
    from xva_trainer_b200 import ops
    B, T, K, N = 2, 70, 64, 384
    x, w = gen(B, T, K, seed=11), gen(1, N, K, seed=12) * K ** -0.5
    res = torch.zeros(B, T, N, device="cuda")
    p, seed = 0.2, 4242
    plain = ops.conv_fwd(x, w, residual=res)
    dropped = ops.conv_fwd(x, w, residual=res, drop_p=p, seed=seed)
    fwd_keep = dropped.abs() > 0
    gamma = torch.ones(N, device="cuda")
    y, sv = ops.layernorm_fwd(plain, gamma, torch.zeros(N, device="cuda"), None)
    dg, db = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    dx, dxd = ops.layernorm_bwd(gen(B, T, N, seed=13), sv, gamma, None, dg, db, want_drop=True, drop_pre_p=p, seed_pre=seed)
    bwd_keep = dxd.abs() > 0
    sel = (plain.abs() > 1e-6) & (dx.abs() > 1e-9)
    assert torch.equal(fwd_keep[sel], bwd_keep[sel])
    assert abs(float(fwd_keep[sel].float().mean()) - (1 - p)) < 0.02


@pytest.mark.parametrize("rows,C,ld", [(5000, 1536, 1536), (777, 192, 192), (300, 80, 96), (100, 1, 32), (64, 30, 30)])
def test_colsum(lib, rows, C, ld):
    from xva_trainer_b200 import ops
    x = gen(rows, ld, seed=21)
    out = torch.ones(C, device="cuda")
    ops.colsum_(rows, C, ld, x, out)
    assert rel(out, 1 + x[:, :C].double().sum(0).float()) < 1e-5


@pytest.mark.parametrize("Z,T,p", [(3, 200, 0.0), (2, 880, 0.1), (4, 160, 0.25)])
def test_fused_softmax_backward_matches_unfused(lib, Z, T, p):
    """XVA_GEMM_SOFTMAX_BWD (dS computed in the epilogue of dP = dO.V^T, row term dO.O) vs the two-kernel path
    (xva_gemm + xva_softmax_bwd) and vs torch autograd through softmax (no dropout case). Same dropout mask by
    construction (same seed and element index); tolerance: tf32 products."""
    import math
    from xva_trainer_b200 import ops
    d = 64
    alpha = 1.0 / math.sqrt(d)
    q, k, v, dvec = (gen(Z, T, d, seed=s) for s in (31, 32, 33, 34))
    lens = torch.tensor([T, max(1, T // 2), T - 3, T][:Z], device="cuda", dtype=torch.int32)
    ld = (T + 31) // 32 * 32
    S = torch.empty(Z, T, ld, device="cuda")
    ops.bmm_nt(q, k, alpha=alpha, out=S[..., :T])
    seed = 777
    P, Pd = ops.softmax_fwd(S, lens, T, p, seed)
    vec = ops.bmm_nn(Pd[..., :T], v)
    # two-kernel path
    dP = torch.empty(Z, T, ld, device="cuda")
    ops.bmm_nt(dvec, v, out=dP[..., :T])
    ops.softmax_bwd_(P, dP, T, alpha, p, seed)
    # fused
    dS = torch.full((Z, T, ld), float("nan"), device="cuda")
    dS[..., T:].zero_()
    D = ops.rowdot2(dvec, vec)
    ops.bmm_nt(dvec, v, alpha=alpha, out=dS[..., :T], round_out=True, softmax_bwd=(P, D, p, seed, None))
    assert rel(dS[..., :T], dP[..., :T]) < 3e-3, rel(dS[..., :T], dP[..., :T])
    dS_ref = torch.empty(Z, T, ld, device="cuda")
    ops.bmm_nt(dvec, v, alpha=alpha, out=dS_ref[..., :T], round_out=True, softmax_bwd=(P, D, p, seed, None), ref=True)
    assert rel(dS[..., :T], dS_ref[..., :T]) < 2e-3
    if p == 0.0:
        torch.backends.cuda.matmul.allow_tf32 = False
        ql = q.clone().requires_grad_(True)
        mask = torch.arange(T, device="cuda")[None, None, :] >= lens[:, None, None]
        s = (ql @ k.transpose(1, 2)) * alpha
        pr = torch.softmax(s.masked_fill(mask, float("-inf")), -1)
        ((pr @ v) * dvec).sum().backward()
        dq = ops.bmm_nn(dS[..., :T], k)
        assert rel(dq, ql.grad) < 3e-3, rel(dq, ql.grad)
