"""CPU stand-in for libxva_b200's C ABI. TEST INFRASTRUCTURE ONLY (like oracle/): nothing in the product package may
import it (tests/test_abi.py greps for it), and the product path still fails loudly without the CUDA library.

Why it exists: the Python host side of the engine (which GEMM is launched with which operands, strides, taps, epilogue
flags, in which order; what the backward saves and re-reads; how parameters are packed and un-packed) is most of what can
go wrong in a module, and it can be checked without a GPU if every C-ABI call it makes is executed on host memory according
to the contract written in include/xva_b200.h. This module does that for EVERY entry point that launches work (70 of the
77; the rest are introspection calls -- tests/test_cabi_emu_cpu.py::test_every_compute_entry_point_of_the_abi_has_a_stand_in):

    xva_gemm / xva_gemm_ref     the tap-GEMM contract exactly as csrc/gemm_ref.cu states it: modes 0 / 1 / 2, taps, row
                                shifts, per-tap column offsets, batched / per-tap B, grouped convolutions, every epilogue
                                step in order (bias, (leaky) ReLU, gate, dropout, residual, LayerNorm, tanh, length mask,
                                second output, the fused softmax backward)
    row / element-wise kernels  softmax, LayerNorm, column sums, embeddings, scalar convolution, row dots, length regulator,
                                average_pitch, masked MSE losses, LAMB, AdamW, gated activation, posterior sample, mel
                                pieces (reflect pad, magnitude, log clamp), GAN losses, first discriminator convolution,
                                pooling -- restated from csrc/rowops.cu, regulate.cu, loss_optim.cu, elemwise.cu, vits.cu,
                                melspec.cu, disc.cu
    xva_wn_pack_* / xva_sn_pack_*  weight-norm / spectral-norm re-parametrisation + packing from their descriptor tables
    xva_attn_fwd / _bwd         the fused attention, from its contract in include/xva_b200.h
    xva_attn_score_*, xva_attn_bin_loss, xva_attn_grad_combine, xva_vits_logp_operands, xva_vits_kl
                                restated from csrc/align.cu / vits.cu
    xva_attn_ctc                NOT the kernel's recursion: torch.nn.functional.ctc_loss + autograd per utterance, as the
                                reference computes it -- an independent implementation of the same contract
    xva_mas_width1              the oracle's own searches (oracle.fastpitch.b_mas, oracle.vits.maximum_path): on the CPU
                                these calls check the host code around the kernel only; the kernel is compared with the
                                same functions bit for bit in tests/test_mas_gpu.py
    xva_text_embed_fwd / _bwd, xva_rel_band_add, xva_rel_band_gather, xva_pad_cols
                                NOT restated: csrc/relattn_body.h (the per-element functions the CUDA kernels loop over)
                                compiled for the host with g++ (tests/relattn_host.cpp) and run as is

Arithmetic is fp32 inputs with float64 accumulation and no tf32 operand rounding (the library's
xva_set_operand_rounding(0) test mode); dropout uses the library's counter hash (csrc/common.cuh), so forward and
backward masks can be checked for consistency.

The emulator is validated in tests/test_cabi_emu_cpu.py by running host code whose GPU parity is established against the
oracle / the reference recordings: whole FastPitch training steps of all four stages (forward, loss, backward, clip + LAMB),
the HiFi-GAN generator and one whole HiFi-GAN training step, the WaveNet stack, the normalising flow, the alignment block
and one whole xVAPitch --hifi_only step. New host code is then checked the same way BEFORE it costs GPU time
(tests/test_vits_text_encoder_cpu.py: the text encoder's first hardware run passed 12 of 12).

Usage:  with cabi_emu.installed():  ... build product modules on device="cpu" via cabi_emu.load_module(...)
"""
import contextlib
import ctypes as C
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, "_emu_build")
SEED_STEP = 0xA24BAED4963EE407
MASK64 = (1 << 64) - 1


# ------------------------------------------------------------------------------------------------ raw memory views
def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, C.c_void_p):
        return p.value or 0
    raise TypeError(f"not a pointer argument: {p!r}")


_CT = {np.float32: C.c_float, np.int32: C.c_int32, np.int64: C.c_int64, np.float64: C.c_double, np.uint64: C.c_uint64}


def flat(p, n, dtype=np.float32):
    """n elements of `dtype` at address p as a writable numpy array (no copy); None for a null pointer."""
    a = _addr(p)
    if not a:
        return None
    return np.ctypeslib.as_array(C.cast(a, C.POINTER(_CT[dtype])), shape=(int(n),))


def strided(p, shape, strides, dtype=np.float32):
    """View of `shape` with ELEMENT strides `strides` at address p."""
    ext = 1 + sum((s - 1) * st for s, st in zip(shape, strides))
    base = flat(p, ext, dtype)
    item = base.itemsize
    return np.lib.stride_tricks.as_strided(base, shape=tuple(int(s) for s in shape), strides=tuple(int(st) * item for st in strides))


# ------------------------------------------------------------------------------------------------ tf32 operand model
# TF32 = False (default): exact arithmetic, the library's xva_set_operand_rounding(0) test mode.
# TF32 = True: the PRODUCT path's operand precision is modelled -- xva_gemm truncates its fp32 operands to tf32 (10-bit
# mantissa) as the tensor cores do, and every kernel that stores a GEMM operand rounds it to nearest (csrc/common.cuh tf32_rn).
# Supported for the entry points in TF32_MODELLED (the text encoder / pitch predictor / FFT-block path); any other call
# raises in this mode rather than silently computing in a precision the device would not use. xva_gemm_ref stays exact.
TF32 = False
TF32_MODELLED = {"xva_gemm", "xva_gemm_ref", "xva_softmax_fwd", "xva_softmax_bwd", "xva_layernorm_fwd", "xva_layernorm_bwd",
                 "xva_colsum", "xva_colsum_items", "xva_round_tf32", "xva_counter_add", "xva_rowdot2", "xva_device_check",
                 "xva_set_operand_rounding", "xva_text_embed_fwd", "xva_text_embed_bwd", "xva_rel_band_add", "xva_rel_band_gather",
                 "xva_pad_cols", "xva_adamw_step",
                 # the FastPitch stages 2-4 path
                 "xva_embed_pos", "xva_embed_bwd", "xva_scalar_conv_add", "xva_scalar_conv_bwd", "xva_rowdot_fwd", "xva_rowdot_bwd",
                 "xva_regulate_len_scan", "xva_regulate_len_fwd", "xva_regulate_len_bwd", "xva_average_pitch", "xva_mel_mse",
                 "xva_mel_mse_grad", "xva_lens_mse", "xva_lens_mse_grad", "xva_grad_sqnorm", "xva_lamb_step", "xva_attn_fwd",
                 "xva_attn_bwd",
                 # the HiFi-GAN / xVAPitch --hifi_only path
                 "xva_wn_pack_fwd", "xva_wn_pack_bwd", "xva_sn_pack_fwd", "xva_sn_pack_bwd", "xva_mean3_lrelu", "xva_sum3",
                 "xva_tanh_bwd", "xva_gated_act_fwd", "xva_gated_act_bwd", "xva_vits_sample_fwd", "xva_vits_sample_bwd",
                 "xva_conv_c1_fwd", "xva_conv_c1_bwd_w", "xva_conv_c1_bwd_x", "xva_avgpool4_fwd", "xva_avgpool4_bwd",
                 "xva_zero_tail_rows", "xva_reflect_pad_fwd", "xva_reflect_pad_bwd", "xva_spec_mag_fwd", "xva_spec_mag_bwd",
                 "xva_log_clamp_fwd", "xva_log_clamp_bwd", "xva_reduce_loss", "xva_loss_grad", "xva_l1_loss_grad"}


def tf32_rn(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, low 13 mantissa bits cleared."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


# What-if model for a 16-bit operand mode (tcgen05 kind::f16 at twice the MMA rate): with
# OPERAND16 = "fp16" / "bf16" (and TF32 = True) producers store, and xva_gemm reads, operands in that type instead of tf32.
# OPERAND_STATS, when a dict, collects per xva_gemm launch the magnitude range of both operands and the fraction of non-zero
# elements below fp16's smallest normal / subnormal -- the numbers a loss scale has to fix.
OPERAND16 = None
OPERAND_STATS = None


def to16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if OPERAND16 == "fp16":
        with np.errstate(over="ignore"):
            return x.astype(np.float16).astype(np.float32)
    u = x.view(np.uint32)                                # bf16: round to nearest even on the upper 16 bits
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


# Ablation switch for the operand model: which producers round. PHASE is set by the caller ("fwd" / "bwd") around the
# forward and backward passes; ROUND_WHAT names the classes that round: "weights" (the operand copy of the parameters:
# xva_round_tf32, the optimizers' p_tf32, the weight packers), "fwd" (operands produced by the forward pass), "bwd" (by the backward).
PHASE = "fwd"
ROUND_WHAT = {"weights", "fwd", "bwd"}
MMA_TRUNCATES = True          # False (ablation only): the MMA reads un-rounded operands exactly


ROUNDING_ON = True            # the library's xva_set_operand_rounding switch (the tests' exact-arithmetic mode turns it off)


def _set_operand_rounding(on):
    global ROUNDING_ON
    ROUNDING_ON = bool(on)
    return 0


def _rn(x, what=None):
    if not TF32 or not ROUNDING_ON or (what or PHASE) not in ROUND_WHAT:
        return x
    return to16(x) if OPERAND16 else tf32_rn(x)


def _operand(x, which, g):
    """The value the MMA reads for an fp32 operand element."""
    if OPERAND_STATS is not None:
        a = np.abs(x[x != 0])
        if a.size:
            key = (g.mode, which)
            st = OPERAND_STATS.setdefault(key, {"n": 0, "min": np.inf, "max": 0.0, "below_normal": 0, "below_subnormal": 0, "over": 0})
            st["n"] += a.size
            st["min"], st["max"] = min(st["min"], float(a.min())), max(st["max"], float(a.max()))
            st["below_normal"] += int((a < 6.1035e-5).sum())
            st["below_subnormal"] += int((a < 5.96e-8).sum())
            st["over"] += int((a > 65504.0).sum())
    return to16(x) if OPERAND16 else (tf32_trunc(x) if MMA_TRUNCATES else x)


# ------------------------------------------------------------------------------------------------ dropout hash
def _hash_u64(seed, idx4):
    with np.errstate(over="ignore"):
        x = idx4 * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & MASK64)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
    return x


def dropout_scale(seed, seed_dev, idx, p):
    """csrc/common.cuh dropout_scale for an array of element indices: 0 where dropped, 1 / (1 - p) where kept."""
    p = float(np.float32(p))
    if p <= 0.0:
        return np.ones(idx.shape, np.float32)
    sd = flat(seed_dev, 1, np.uint64)
    eff = (int(seed) + (int(sd[0]) * SEED_STEP if sd is not None else 0)) & MASK64
    thresh = int(p * 4294967296.0) & 0xFFFFFFFF
    idx = idx.astype(np.uint64)
    h = _hash_u64(eff, idx >> np.uint64(2))
    f = (h >> (np.uint64(16) * (idx & np.uint64(3)))) & np.uint64(0xFFFF)
    keep = f >= np.uint64(thresh >> 16)
    return np.where(keep, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ tap-GEMM
from_flags = dict(RELU=1 << 0, LN=1 << 1, DROP_PRE=1 << 2, DROP_POST=1 << 3, ATOMIC=1 << 4, LRELU_GATE=1 << 5, ROUND_OUT=1 << 6,
                  TANH=1 << 7, SOFTMAX_BWD=1 << 8, HALO=1 << 9)


def _gemm_grouped(g, G):
    """xva_gemm_args.groups > 1 (DiscriminatorS's grouped convolutions in one launch), csrc/gemm_ref.cu ref_dot / wgrad:
       mode 0: out[.., n] contracts A columns a_col[j] + grp * grp_step + [0, K) with B row n,            grp = n // (N / G)
       mode 1: out[.., n] contracts A columns a_col[j] + grp * K + [0, K) with B rows grp * K + [0, K), column n - grp * N / G
       mode 2: out[j, m, n] contracts A column m with B column a_col[j] + grp * grp_step + n,              grp = m // (M / G)
    Epilogue (mode 0 / 1) as in the dense case without LayerNorm / softmax-backward."""
    Z, R, N, K, taps = g.Z, g.R, g.N, g.K, g.taps
    F = from_flags
    assert not (g.flags & (F["LN"] | F["SOFTMAX_BWD"])) and g.b_batch_z == 0
    if g.mode in (0, 1):
        a_rows = g.a_rows or R
        n_per = N // G
        step = g.grp_step if g.mode == 0 else K
        a_cols = max(g.a_col[j] for j in range(taps)) + (G - 1) * step + K
        assert Z == 1 or g.a_zs != 0
        A = strided(g.a, (Z, a_rows, a_cols), (g.a_zs, g.a_rs, 1))
        n_zb = (taps - 1) * g.b_tap_z + 1
        if g.mode == 0:
            Bm = strided(g.b, (n_zb, N, K), (g.b_zs, g.b_rs, 1))
        else:
            Bm = strided(g.b, (n_zb, G * K, n_per), (g.b_zs, g.b_rs, 1))
        acc = np.zeros((Z, R, N), np.float64)
        for j in range(taps):
            rr = np.arange(R) + g.shift[j]
            ok = (rr >= 0) & (rr < a_rows)
            if not ok.any():
                continue
            Bj = Bm[j * g.b_tap_z]
            Bj = (_operand(np.ascontiguousarray(Bj), "B", g) if TF32 else Bj).astype(np.float64)
            for grp in range(G):
                Aj = np.zeros((Z, R, K), np.float32)
                c0 = g.a_col[j] + grp * step
                Aj[:, ok] = A[:, rr[ok], c0:c0 + K]
                Aj = (_operand(Aj, "A", g) if TF32 else Aj).astype(np.float64)
                if g.mode == 0:
                    acc[:, :, grp * n_per:(grp + 1) * n_per] += Aj @ Bj[grp * n_per:(grp + 1) * n_per].T
                else:
                    acc[:, :, grp * n_per:(grp + 1) * n_per] += Aj @ Bj[grp * K:(grp + 1) * K]
        v = (np.float64(np.float32(g.alpha)) * acc).astype(np.float32)
        view = lambda p, rs, zs: strided(p, (Z, R, N), (zs, rs, 1))
        if g.bias:
            v = v + flat(g.bias, N)[None, None, :]
        if g.flags & F["RELU"]:
            v = np.where(v > 0, v, np.float32(g.act_slope) * v)
        if g.gate:
            v = v * np.where(view(g.gate, g.g_rs, g.g_zs) > 0, np.float32(1.0), np.float32(g.gate_slope))
        assert not ((g.flags & F["DROP_PRE"]) and g.drop_p > 0)
        if g.residual:
            v = v + view(g.residual, g.r_rs, g.r_zs)
        if g.flags & F["TANH"]:
            v = np.tanh(v)
        if g.lens:
            lens = flat(g.lens, Z, np.int32)
            v = v * (np.arange(R)[None, :] < lens[:, None]).astype(np.float32)[:, :, None]
        v = v.astype(np.float32)
        if g.out_act:
            view(g.out_act, g.o_rs, g.o_zs)[...] = _rn(np.where(v > 0, v, np.float32(g.out_act_slope) * v).astype(np.float32))
        view(g.out, g.o_rs, g.o_zs)[...] = _rn(v) if (g.flags & F["ROUND_OUT"]) else v
        return 0
    M, ZR = g.M, g.ZR
    a_rows, b_rows = g.a_rows or R, g.b_rows or R
    og = M // G
    b_cols = max(g.a_col[j] for j in range(taps)) + (G - 1) * g.grp_step + N
    tA = min(R, a_rows)
    A = strided(g.a, (Z, tA, M), (g.a_zs, g.a_rs, 1)).astype(np.float64)
    Bm = strided(g.b, (Z, b_rows, b_cols), (g.b_zs, g.b_rs, 1))
    for zo in range(Z // ZR):
        for j in range(taps):
            bt = np.arange(tA) + g.shift[j]
            ok = (bt >= 0) & (bt < b_rows)
            acc = np.zeros((M, N), np.float64)
            if ok.any():
                for zr in range(ZR):
                    z = zo * ZR + zr
                    for grp in range(G):
                        c0 = g.a_col[j] + grp * g.grp_step
                        At = np.ascontiguousarray(A[z, ok][:, grp * og:(grp + 1) * og].T, dtype=np.float32)
                        Bt = np.ascontiguousarray(Bm[z, bt[ok], c0:c0 + N], dtype=np.float32)
                        if TF32:
                            At, Bt = _operand(At, "A", g), _operand(Bt, "B", g)
                        acc[grp * og:(grp + 1) * og] += At.astype(np.float64) @ Bt.astype(np.float64)
            o = strided(_addr(g.out) + 4 * (zo * g.o_zs + j * g.o_js), (M, N), (g.o_rs, 1))
            val = (np.float64(np.float32(g.alpha)) * acc).astype(np.float32)
            if g.flags & F["ATOMIC"]:
                o += val
            else:
                o[...] = val
    return 0


def _up32(n):
    return (n + 31) // 32 * 32


def check_gemm_args(g):
    """The argument checks of csrc/gemm_tc.cu (gemm_tc_launch / encode_map) that do not depend on the tile plan: a launch
    the library would refuse with XVA_ERR_ARG must not pass on the CPU either. TMA needs 16-byte aligned operand bases and
    row / item pitches; MN-major operands are fetched in 32-column chunks."""
    assert 0 <= g.mode <= 2 and 1 <= g.taps <= 48 and g.Z >= 1 and g.R >= 1 and g.N >= 1, "gemm: empty / bad problem"
    assert g.a and g.b and g.out, "gemm: null operand"
    for name, ptr, rs, zs, nz in (("a", g.a, g.a_rs, g.a_zs, g.Z), ("b", g.b, g.b_rs, g.b_zs, g.b_nz if g.mode != 2 else g.Z)):
        assert ptr % 16 == 0, f"gemm: operand {name} base {ptr:#x} is not 16-byte aligned (TMA)"
        assert (rs * 4) % 16 == 0, f"gemm: operand {name} row pitch {rs} floats is not a multiple of 16 bytes"
        assert nz <= 1 or zs == 0 or (zs * 4) % 16 == 0, f"gemm: operand {name} item pitch {zs} floats is not a multiple of 16 bytes"
    G = g.groups if g.groups > 1 else 1
    if g.mode != 2:
        assert g.K >= 1
        if g.flags & from_flags["LN"]:
            assert g.N % 16 == 0 and g.N <= 512 and g.gamma and g.beta, f"gemm: LayerNorm epilogue needs N % 16 == 0, <= 512 (N={g.N})"
        if g.mode == 1:
            assert g.N % 32 == 0 or g.b_rs >= _up32(g.N), f"gemm: MN-major B with N={g.N} needs N % 32 == 0 or a row stride >= {_up32(g.N)} (got {g.b_rs})"
        if G > 1:
            assert not (g.flags & from_flags["LN"]) and g.b_batch_z == 0 and g.N % G == 0
            if g.mode == 0:
                assert (g.N // G) % 16 == 0 and g.N // G <= 256
            else:
                assert (g.N // G) % 32 == 0 and g.N // G <= 256 and g.K % 32 == 0
    else:
        assert g.M >= 1 and g.ZR >= 1 and g.Z % g.ZR == 0
        assert g.M % 32 == 0 or g.a_rs >= _up32(g.M), f"gemm: MN-major A with M={g.M} needs M % 32 == 0 or a row stride >= {_up32(g.M)} (got {g.a_rs})"
        assert g.N % 32 == 0 or g.b_rs >= _up32(g.N), f"gemm: MN-major B with N={g.N} needs N % 32 == 0 or a row stride >= {_up32(g.N)} (got {g.b_rs})"
        for j in range(g.taps):
            assert g.a_col[j] % 32 == 0, f"gemm: wgrad column offset {g.a_col[j]} of tap {j} is not a multiple of 32"
        assert g.split <= 1 or (g.flags & from_flags["ATOMIC"]), "gemm: split > 1 needs GEMM_ATOMIC"
        if G > 1:
            og = g.M // G
            assert g.M % G == 0 and og in (32, 64, 128) and g.grp_step == g.N and g.N % 32 == 0 and (128 // og) * g.N <= 256


def _gemm_ref(ref, stream=None):
    return _gemm(ref, stream, exact=True)


def _gemm(ref, stream=None, exact=False):
    g = ref._obj
    check_gemm_args(g)
    tf = TF32 and not exact
    Z, R, N, K, taps = g.Z, g.R, g.N, g.K, g.taps
    F = from_flags
    G = g.groups if g.groups > 1 else 1
    if G > 1:
        return _gemm_grouped(g, G)
    if g.mode in (0, 1):
        a_rows = g.a_rows or R
        a_cols = max(g.a_col[j] for j in range(taps)) + K
        assert Z == 1 or g.a_zs != 0, "cabi_emu: a_zs == 0 with Z > 1 means different things to xva_gemm and xva_gemm_ref"
        A = strided(g.a, (Z, a_rows, a_cols), (g.a_zs, g.a_rs, 1))
        n_zb = max(j * g.b_tap_z + z * g.b_batch_z for j in range(taps) for z in (0, Z - 1)) + 1
        if g.mode == 0:
            nv = min(N, g.b_rows) if g.b_rows else N
            Bm = strided(g.b, (n_zb, nv, K), (g.b_zs, g.b_rs, 1))
        else:
            kv = min(K, g.b_rows) if g.b_rows else K
            Bm = strided(g.b, (n_zb, kv, N), (g.b_zs, g.b_rs, 1))
        # products in fp32 BLAS (what the exact checker kernel computes with fmaf), taps summed in float64
        acc = np.zeros((Z, R, N), np.float64)
        for j in range(taps):
            rr = np.arange(R) + g.shift[j]
            ok = (rr >= 0) & (rr < a_rows)
            if not ok.any():
                continue
            Aj = np.zeros((Z, R, K), np.float32)
            Aj[:, ok] = A[:, rr[ok], g.a_col[j]:g.a_col[j] + K]
            if tf:
                Aj = _operand(Aj, "A", g)
            if g.b_batch_z == 0:
                Bz = np.ascontiguousarray(Bm[j * g.b_tap_z])
                if tf:
                    Bz = _operand(Bz, "B", g)
                if g.mode == 0:
                    acc[:, :, :Bz.shape[0]] += (Aj.reshape(Z * R, K) @ Bz.T).reshape(Z, R, -1)
                else:
                    acc += (Aj.reshape(Z * R, K)[:, :Bz.shape[0]] @ Bz).reshape(Z, R, N)
                continue
            for z in range(Z):
                Bz = Bm[j * g.b_tap_z + z * g.b_batch_z]
                if tf:
                    Bz = _operand(Bz, "B", g)
                if g.mode == 0:
                    acc[z, :, :Bz.shape[0]] += Aj[z] @ Bz.T
                else:
                    acc[z] += Aj[z, :, :Bz.shape[0]] @ Bz
        rows_idx = (np.arange(Z)[:, None] * R + np.arange(R)[None, :]).astype(np.uint64)           # z * R + r
        view = lambda p, rs, zs: strided(p, (Z, R, N), (zs, rs, 1))
        if g.flags & F["SOFTMAX_BWD"]:
            ld = g.drop_ld if g.drop_ld > 0 else N
            di = rows_idx[:, :, None] * np.uint64(ld) + np.arange(N, dtype=np.uint64)[None, None, :]
            dp = acc.astype(np.float32) * dropout_scale(g.seed, g.seed_dev, di, g.drop_p)
            rowvec = flat(g.rowvec, Z * R).reshape(Z, R, 1)
            v = np.float32(g.alpha) * view(g.gate, g.g_rs, g.g_zs) * (dp - rowvec)
        else:
            v = (np.float64(np.float32(g.alpha)) * acc).astype(np.float32)
            if g.bias:
                v = v + flat(g.bias, N)[None, None, :]
            if g.flags & F["RELU"]:
                v = np.where(v > 0, v, np.float32(g.act_slope) * v)
            if g.gate:
                v = v * np.where(view(g.gate, g.g_rs, g.g_zs) > 0, np.float32(1.0), np.float32(g.gate_slope))
            if (g.flags & F["DROP_PRE"]) and g.drop_p > 0:
                di = rows_idx[:, :, None] * np.uint64(N) + np.arange(N, dtype=np.uint64)[None, None, :]
                v = v * dropout_scale(g.seed, g.seed_dev, di, g.drop_p)
            if g.residual:
                v = v + view(g.residual, g.r_rs, g.r_zs)
        v = v.astype(np.float32)
        keep_row = np.ones((Z, R, 1), np.float32)
        if g.lens:
            lens = flat(g.lens, Z, np.int32)
            keep_row = (np.arange(R)[None, :] < lens[:, None]).astype(np.float32)[:, :, None]
        out = view(g.out, g.o_rs, g.o_zs)
        if not (g.flags & F["LN"]):
            if g.flags & F["TANH"]:
                v = np.tanh(v)
            v = v * keep_row
            if g.out_act:
                view(g.out_act, g.o_rs, g.o_zs)[...] = _rn(np.where(v > 0, v, np.float32(g.out_act_slope) * v))
            out[...] = _rn(v) if (g.flags & F["ROUND_OUT"]) else v
            return 0
        mean = v.astype(np.float64).mean(axis=2, keepdims=True)
        var = ((v - mean) ** 2).mean(axis=2, keepdims=True)
        rstd = 1.0 / np.sqrt(var + np.float32(g.ln_eps))
        y = ((v - mean) * rstd * flat(g.gamma, N)[None, None, :] + flat(g.beta, N)[None, None, :]).astype(np.float32)
        if (g.flags & F["DROP_POST"]) and g.drop_p > 0:
            di = rows_idx[:, :, None] * np.uint64(N) + np.arange(N, dtype=np.uint64)[None, None, :]
            y = y * dropout_scale(g.seed, g.seed_dev, di, g.drop_p)
        out[...] = _rn(y * keep_row) if (g.flags & F["ROUND_OUT"]) else y * keep_row
        if g.out_pre:
            view(g.out_pre, g.o_rs, g.o_zs)[...] = v
        if g.ln_mean:
            flat(g.ln_mean, Z * R)[...] = mean.reshape(-1).astype(np.float32)
        if g.ln_rstd:
            flat(g.ln_rstd, Z * R)[...] = rstd.reshape(-1).astype(np.float32)
        return 0
    # ---- mode 2
    M, ZR = g.M, g.ZR
    a_rows = g.a_rows or R
    b_rows = g.b_rows or R
    b_cols = max(g.a_col[j] for j in range(taps)) + N
    tA = min(R, a_rows)
    assert Z == 1 or (g.a_zs != 0 and g.b_zs != 0)
    A = strided(g.a, (Z, tA, M), (g.a_zs, g.a_rs, 1))
    Bm = strided(g.b, (Z, b_rows, b_cols), (g.b_zs, g.b_rs, 1))
    for zo in range(Z // ZR):
        for j in range(taps):
            acc = np.zeros((M, N), np.float64)
            bt = np.arange(tA) + g.shift[j]
            ok = (bt >= 0) & (bt < b_rows)
            if ok.any():
                for zr in range(ZR):
                    z = zo * ZR + zr
                    At, Bt = np.ascontiguousarray(A[z, ok].T), np.ascontiguousarray(Bm[z, bt[ok], g.a_col[j]:g.a_col[j] + N])
                    acc += (_operand(At, "A", g) @ _operand(Bt, "B", g)) if tf else (At @ Bt)
            o = strided(_addr(g.out) + 4 * (zo * g.o_zs + j * g.o_js), (M, N), (g.o_rs, 1))
            val = (np.float64(np.float32(g.alpha)) * acc).astype(np.float32)
            if g.flags & F["ATOMIC"]:
                o += val
            else:
                o[...] = _rn(val) if (g.flags & F["ROUND_OUT"]) else val
    return 0


# ------------------------------------------------------------------------------------------------ row kernels
def _softmax_fwd(s, lens, Z, R, N, ld, p_out, pd_out, drop_p, seed, seed_dev, stream=None):
    ld = ld if ld > 0 else N
    assert N >= 1 and N <= ld <= 1024, f"softmax: N={N} ld={ld} out of range (max 1024)"
    rows = Z * R
    S = flat(s, rows * ld).reshape(rows, ld)
    nk = np.full(rows, N)
    if _addr(lens):
        nk = np.minimum(np.repeat(flat(lens, Z, np.int32), R), N)
    assert (nk > 0).all(), "cabi_emu: softmax row without a key"
    live = np.arange(ld)[None, :] < nk[:, None]
    v = np.where(live, S, -np.inf).astype(np.float64)
    v = np.exp(v - v.max(axis=1, keepdims=True)) * live
    p = (v / v.sum(axis=1, keepdims=True)).astype(np.float32)
    p = _rn(p)                                           # both outputs are GEMM operands (P.V, dV = P^T dO)
    flat(p_out, rows * ld).reshape(rows, ld)[...] = p
    if _addr(pd_out):
        idx = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(ld) + np.arange(ld, dtype=np.uint64)[None, :]
        flat(pd_out, rows * ld).reshape(rows, ld)[...] = _rn(p * dropout_scale(seed, seed_dev, idx, drop_p))
    return 0


def _softmax_bwd(p, dpd, Z, R, N, ld, alpha, drop_p, seed, seed_dev, stream=None):
    ld = ld if ld > 0 else N
    rows = Z * R
    P = flat(p, rows * ld).reshape(rows, ld)[:, :N].astype(np.float64)
    D = flat(dpd, rows * ld).reshape(rows, ld)
    idx = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(ld) + np.arange(N, dtype=np.uint64)[None, :]
    gv = D[:, :N].astype(np.float64) * dropout_scale(seed, seed_dev, idx, drop_p)
    dot = (P * gv).sum(axis=1, keepdims=True)
    D[:, :N] = _rn((np.float32(alpha) * P * (gv - dot)).astype(np.float32))
    D[:, N:] = 0.0
    return 0


def _layernorm_fwd(x, gamma, beta, lens, Z, R, Cc, eps, y, mean, rstd, stream=None):
    assert 4 <= Cc <= 1024 and Cc % 4 == 0, f"layernorm fwd: C={Cc} (multiple of 4, max 1024)"
    assert all(_addr(q) % 16 == 0 for q in (x, y, gamma, beta)), "layernorm fwd: pointers must be 16-byte aligned"
    rows = Z * R
    X = flat(x, rows * Cc).reshape(rows, Cc).astype(np.float64)
    mu = X.mean(axis=1, keepdims=True)
    rs = 1.0 / np.sqrt(((X - mu) ** 2).mean(axis=1, keepdims=True) + np.float32(eps))
    live = np.ones((rows, 1))
    if _addr(lens):
        live = (np.tile(np.arange(R), Z) < np.repeat(flat(lens, Z, np.int32), R)).astype(np.float64)[:, None]
    flat(y, rows * Cc).reshape(rows, Cc)[...] = _rn((((X - mu) * rs * flat(gamma, Cc) + flat(beta, Cc)) * live).astype(np.float32))
    flat(mean, rows)[...] = mu[:, 0].astype(np.float32)
    flat(rstd, rows)[...] = rs[:, 0].astype(np.float32)
    return 0


def _layernorm_bwd(dy, x, mean, rstd, gamma, lens, Z, R, Cc, dx, dx_drop, dgamma, dbeta, dbias, drop_post_p, seed_post,
                   drop_pre_p, seed_pre, seed_dev, relu_gate, stream=None):
    assert 1 <= Cc <= 1024, f"layernorm bwd: C={Cc} (max 1024)"
    rows = Z * R
    live = np.ones((rows, 1))
    if _addr(lens):
        live = (np.tile(np.arange(R), Z) < np.repeat(flat(lens, Z, np.int32), R)).astype(np.float64)[:, None]
    idx = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(Cc) + np.arange(Cc, dtype=np.uint64)[None, :]
    X = flat(x, rows * Cc).reshape(rows, Cc).astype(np.float64)
    d = flat(dy, rows * Cc).reshape(rows, Cc).astype(np.float64) * dropout_scale(seed_post, seed_dev, idx, drop_post_p) * live
    mu, rs = flat(mean, rows).astype(np.float64)[:, None], flat(rstd, rows).astype(np.float64)[:, None]
    xh = (X - mu) * rs * live
    gm = flat(gamma, Cc).astype(np.float64)[None, :]
    g = d * gm
    s1, s2 = g.mean(axis=1, keepdims=True), (g * xh).mean(axis=1, keepdims=True)
    v = rs * (g - s1 - xh * s2)
    if relu_gate:
        v = np.where((X > 0) & (live > 0), v, 0.0)
    if _addr(dgamma):
        flat(dgamma, Cc)[...] += (d * xh).sum(axis=0).astype(np.float32)
    if _addr(dbeta):
        flat(dbeta, Cc)[...] += d.sum(axis=0).astype(np.float32)
    # the tensor that feeds the dgrad / wgrad GEMMs (dx_drop if there is one, else dx) is stored tf32-rounded
    branch = v
    if _addr(dx_drop):
        flat(dx, rows * Cc).reshape(rows, Cc)[...] = v.astype(np.float32)
        branch = _rn((v * dropout_scale(seed_pre, seed_dev, idx, drop_pre_p)).astype(np.float32)).astype(np.float64)
        flat(dx_drop, rows * Cc).reshape(rows, Cc)[...] = branch.astype(np.float32)
    else:
        branch = _rn(v.astype(np.float32)).astype(np.float64)
        flat(dx, rows * Cc).reshape(rows, Cc)[...] = branch.astype(np.float32)
    if _addr(dbias):
        flat(dbias, Cc)[...] += branch.sum(axis=0).astype(np.float32)
    return 0


def _colsum(x, rows, Cc, ld, out, stream=None):
    if rows:
        flat(out, Cc)[...] += strided(x, (rows, Cc), (ld, 1)).astype(np.float64).sum(axis=0).astype(np.float32)
    return 0


def _colsum_items(x, Z, rows, Cc, ld, zs, out, out_ld, stream=None):
    o = strided(out, (Z, Cc), (out_ld, 1))
    o += strided(x, (Z, rows, Cc), (zs, ld, 1)).astype(np.float64).sum(axis=1).astype(np.float32)
    return 0


def _round_tf32(src, dst, n, stream=None):
    flat(dst, n)[...] = _rn(flat(src, n), "weights")  # (a plain copy in the exact mode)
    return 0


def _counter_add(counter, inc, stream=None):
    with np.errstate(over="ignore"):
        flat(counter, 1, np.uint64)[0] += np.uint64(inc)
    return 0


def _rowdot2(a, b, rows, Cc, a_ld, b_ld, out, stream=None):
    flat(out, rows)[...] = (strided(a, (rows, Cc), (a_ld, 1)).astype(np.float64) *
                            strided(b, (rows, Cc), (b_ld, 1))).sum(axis=1).astype(np.float32)
    return 0


# ------------------------------------------------------------------------------------------------ host build of relattn_body.h
_host_lib = None


def host_lib():
    """relattn_body.h compiled with g++ behind the same extern "C" signatures (tests/relattn_host.cpp)."""
    global _host_lib
    if _host_lib is not None:
        return _host_lib
    src = os.path.join(HERE, "relattn_host.cpp")
    hdr = os.path.join(ROOT, "xva-trainer_b200", "csrc", "relattn_body.h")
    so = os.path.join(_BUILD, "librelattn_host.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(_BUILD, exist_ok=True)
        tmp = f"{so}.{os.getpid()}"
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Werror", "-o", tmp, src], check=True, cwd=HERE)
        os.replace(tmp, so)
    lib = C.CDLL(so)
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from xva_trainer_b200 import capi
    for name in HOST_COMPILED:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = capi.PROTOTYPES[name]
    lib.xva_emu_set_rounding.restype, lib.xva_emu_set_rounding.argtypes = None, [C.c_int]
    _host_lib = lib
    return lib


HOST_COMPILED = ("xva_text_embed_fwd", "xva_text_embed_bwd", "xva_rel_band_add", "xva_rel_band_gather", "xva_pad_cols")

TABLE = {
    "xva_gemm": _gemm, "xva_gemm_ref": _gemm_ref, "xva_softmax_fwd": _softmax_fwd, "xva_softmax_bwd": _softmax_bwd,
    "xva_layernorm_fwd": _layernorm_fwd, "xva_layernorm_bwd": _layernorm_bwd, "xva_colsum": _colsum,
    "xva_colsum_items": _colsum_items, "xva_round_tf32": _round_tf32, "xva_counter_add": _counter_add, "xva_rowdot2": _rowdot2,
    "xva_device_check": lambda *a: 0, "xva_set_operand_rounding": _set_operand_rounding,
}

calls = []          # names of the entry points executed since the last reset (the tests assert on coverage)


def call(name, *args):
    calls.append(name)
    if TF32 and name not in TF32_MODELLED:
        raise NotImplementedError(f"cabi_emu: {name} has no tf32 operand model (TF32 mode covers {sorted(TF32_MODELLED)})")
    if name in HOST_COMPILED:
        host_lib().xva_emu_set_rounding(1 if (TF32 and ROUNDING_ON) else 0)
        rc = getattr(host_lib(), name)(*args)
    elif name in TABLE:
        rc = TABLE[name](*args)
    else:
        raise NotImplementedError(f"cabi_emu: {name} is not emulated")
    if rc != 0:
        raise RuntimeError(f"cabi_emu: {name} returned {rc}")


@contextlib.contextmanager
def installed():
    """capi.call / ops' CUDA checks replaced by the emulator for the duration of the block."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from xva_trainer_b200 import capi, ops
    saved = (capi.load, capi.call, ops._stream, ops._check3)
    capi.load = lambda: types.SimpleNamespace(
        xva_attn_ctc_workspace_bytes=lambda B, T, Tt: 2 * B * T * (2 * Tt + 1) * 8 + B * 8 + (B * T * 4 + 7) // 8 * 8)
    capi.call = call
    ops._stream = lambda: None

    def check3(t, name):
        import torch
        if t.dtype != torch.float32 or t.dim() != 3 or t.stride(2) != 1:
            raise ValueError(f"{name}: expected an fp32 [batch, rows, cols] tensor with contiguous last dim")

    ops._check3 = check3
    del calls[:]
    try:
        yield
    finally:
        capi.load, capi.call, ops._stream, ops._check3 = saved


def load_module(modname, patches, extra_modules=None):
    """A private copy of a product module with its refuse-anything-but-CUDA checks patched out (the same device the
    launch-sequence dry run uses, tests/launch_sequence.py)."""
    src = open(os.path.join(ROOT, "xva-trainer_b200", modname + ".py")).read()
    for a, b in patches:
        assert a in src, f"{modname}: patch anchor not found: {a}"
        src = src.replace(a, b)
    m = types.ModuleType(f"xva_trainer_b200.{modname}_emu")
    m.__package__ = "xva_trainer_b200"
    saved = {}
    for k, v in (extra_modules or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        exec(compile(src, modname + "_emu", "exec"), m.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


# ------------------------------------------------------------------------------------------------ FastPitch-only entry points
def _lens_mask(lens, Z, R):
    if not _addr(lens):
        return np.ones((Z, R), bool)
    return np.arange(R)[None, :] < flat(lens, Z, np.int32)[:, None]


def _embed_pos(tokens, emb, inp, lens, inv_freq, B, T, Cc, out, stream=None):
    O = flat(out, B * T * Cc).reshape(B, T, Cc)
    if _addr(tokens):
        tok = flat(tokens, B * T, np.int64).reshape(B, T)
        n_rows = int(tok.max()) + 1
        v = flat(emb, n_rows * Cc).reshape(n_rows, Cc)[tok]
        live = tok != 0
    else:
        v = flat(inp, B * T * Cc).reshape(B, T, Cc).copy()
        live = _lens_mask(lens, B, T)
    if _addr(inv_freq):
        f = flat(inv_freq, Cc // 2)
        ang = (np.arange(T, dtype=np.float32)[:, None] * f[None, :]).astype(np.float32)
        pos = np.concatenate([np.sin(ang), np.cos(ang)], axis=1)
        v = v + pos[None] * live[:, :, None]
    O[...] = _rn(v.astype(np.float32))                  # first GEMM operand of the FFT stack
    return 0


def _embed_bwd(tokens, dout, B, T, Cc, demb, stream=None):
    tok = flat(tokens, B * T, np.int64)
    D = flat(dout, B * T * Cc).reshape(B * T, Cc)
    n_rows = int(tok.max()) + 1
    E = flat(demb, n_rows * Cc).reshape(n_rows, Cc)
    live = tok != 0
    np.add.at(E, tok[live], D[live])
    return 0


def _scalar_conv_add(io, x, w, bias, lens, B, T, Cc, stream=None):
    IO = flat(io, B * T * Cc).reshape(B, T, Cc)
    X = flat(x, B * T).reshape(B, T)
    Wt, bs = flat(w, Cc * 3).reshape(Cc, 3), flat(bias, Cc)
    xp = np.pad(X, ((0, 0), (1, 1)))
    add = bs[None, None, :] + xp[:, :-2, None] * Wt[None, None, :, 0] + X[:, :, None] * Wt[None, None, :, 1] + xp[:, 2:, None] * Wt[None, None, :, 2]
    live = _lens_mask(lens, B, T)
    IO[...] = np.where(live[:, :, None], _rn((IO + add).astype(np.float32)), IO).astype(np.float32)
    return 0


def _scalar_conv_bwd(dout, x, B, T, Cc, dw, dbias, stream=None):
    D = flat(dout, B * T * Cc).reshape(B, T, Cc).astype(np.float64)
    X = flat(x, B * T).reshape(B, T).astype(np.float64)
    xp = np.pad(X, ((0, 0), (1, 1)))
    DW = flat(dw, Cc * 3).reshape(Cc, 3)
    for j, xs in enumerate((xp[:, :-2], X, xp[:, 2:])):
        DW[:, j] += np.einsum("btc,bt->c", D, xs).astype(np.float32)
    flat(dbias, Cc)[...] += D.sum(axis=(0, 1)).astype(np.float32)
    return 0


def _rowdot_fwd(x, w, bias, lens, Z, R, Cc, out, stream=None):
    X = flat(x, Z * R * Cc).reshape(Z, R, Cc).astype(np.float64)
    v = X @ flat(w, Cc).astype(np.float64) + float(flat(bias, 1)[0])
    flat(out, Z * R).reshape(Z, R)[...] = np.where(_lens_mask(lens, Z, R), v, 0.0).astype(np.float32)
    return 0


def _rowdot_bwd(dout, x, w, lens, Z, R, Cc, dx, dw, db, stream=None):
    g = np.where(_lens_mask(lens, Z, R), flat(dout, Z * R).reshape(Z, R), 0.0).astype(np.float64)
    X = flat(x, Z * R * Cc).reshape(Z, R, Cc).astype(np.float64)
    flat(dx, Z * R * Cc).reshape(Z, R, Cc)[...] = (g[:, :, None] * flat(w, Cc)[None, None, :]).astype(np.float32)
    flat(dw, Cc)[...] += np.einsum("zr,zrc->c", g, X).astype(np.float32)
    flat(db, 1)[0] += np.float32(g.sum())
    return 0


def _regulate_scan(durs, B, Tt, pace, mel_max_len, cum, dec_lens, stream=None):
    d = flat(durs, B * Tt).reshape(B, Tt)
    reps = ((d * np.float32(pace)).astype(np.float32) + np.float32(0.5)).astype(np.float32).astype(np.int64)   # trunc toward zero
    Cm = flat(cum, B * (Tt + 1), np.int32).reshape(B, Tt + 1)
    Cm[:, 0] = 0
    Cm[:, 1:] = np.cumsum(reps, axis=1)
    tot = Cm[:, Tt]
    flat(dec_lens, B, np.int32)[...] = np.where((mel_max_len >= 0) & (tot > mel_max_len), mel_max_len, tot)
    return 0


def _token_of_frame(Cm, T_out):
    """idx[b, t] = j with cum[b, j] <= t < cum[b, j + 1], -1 past the total."""
    B = Cm.shape[0]
    idx = np.full((B, T_out), -1, np.int64)
    for b in range(B):
        t = np.arange(T_out)
        j = np.searchsorted(Cm[b], t, side="right") - 1
        ok = t < Cm[b, -1]
        idx[b, ok] = j[ok]
    return idx


def _regulate_fwd(enc, cum, B, Tt, Cc, T_out, out, idx_out, stream=None):
    Cm = flat(cum, B * (Tt + 1), np.int32).reshape(B, Tt + 1)
    E = flat(enc, B * Tt * Cc).reshape(B, Tt, Cc)
    idx = _token_of_frame(Cm, T_out)
    O = flat(out, B * T_out * Cc).reshape(B, T_out, Cc)
    for b in range(B):
        O[b] = np.where((idx[b] >= 0)[:, None], E[b, np.clip(idx[b], 0, Tt - 1)], 0.0)
    if _addr(idx_out):
        flat(idx_out, B * T_out, np.int32).reshape(B, T_out)[...] = idx
    return 0


def _regulate_bwd(dout, cum, B, Tt, Cc, T_out, denc, accumulate, stream=None):
    Cm = flat(cum, B * (Tt + 1), np.int32).reshape(B, Tt + 1)
    D = flat(dout, B * T_out * Cc).reshape(B, T_out, Cc)
    idx = _token_of_frame(Cm, T_out)
    G = flat(denc, B * Tt * Cc).reshape(B, Tt, Cc)
    for b in range(B):
        acc = np.zeros((Tt, Cc), np.float64)
        ok = idx[b] >= 0
        np.add.at(acc, idx[b][ok], D[b][ok])
        G[b] = (G[b] + acc if accumulate else acc).astype(np.float32)
    return 0


def _average_pitch(pitch, durs, B, Fn, Tm, Tt, out, log1p_out, stream=None):
    P = flat(pitch, B * Fn * Tm).reshape(B, Fn, Tm)
    d = flat(durs, B * Tt).reshape(B, Tt)
    O = flat(out, B * Fn * Tt).reshape(B, Fn, Tt)
    for b in range(B):
        run = np.cumsum(d[b], dtype=np.float32)                       # fp32 running sum, truncated like cumsum(...).long()
        cm = np.concatenate([[0], run.astype(np.int64)])
        for j in range(Tt):
            t0, t1 = min(cm[j], Tm), min(cm[j + 1], Tm)
            seg = P[b, :, t0:t1]
            cnt = (seg != 0).sum(axis=1)
            mean = np.where(cnt > 0, seg.sum(axis=1, dtype=np.float32) / np.maximum(cnt, 1), 0.0).astype(np.float32)
            O[b, :, j] = np.log(np.float32(1.0) + mean) if log1p_out else mean
    return 0


def _mel_mse(pred, tgt, B, T_out, Tm, Cc, acc, stream=None):
    Pm = flat(pred, B * T_out * Cc).reshape(B, T_out, Cc)
    Y = flat(tgt, B * Cc * Tm).reshape(B, Cc, Tm).transpose(0, 2, 1)          # [B, Tm, C]
    Pf = np.zeros((B, Tm, Cc), np.float32)
    Pf[:, :min(T_out, Tm)] = Pm[:, :min(T_out, Tm)]
    m = Y != 0
    A = flat(acc, 2, np.float64)
    A[0] += float((((Pf - Y).astype(np.float32) ** 2).astype(np.float64) * m).sum())
    A[1] += float(m.sum())
    return 0


def _mel_mse_grad(pred, tgt, B, T_out, Tm, Cc, ldd, acc, scale, dpred, stream=None):
    Pm = flat(pred, B * T_out * Cc).reshape(B, T_out, Cc)
    Y = flat(tgt, B * Cc * Tm).reshape(B, Cc, Tm).transpose(0, 2, 1)[:, :T_out]
    k = np.float32(2.0 * scale / flat(acc, 2, np.float64)[1])
    D = flat(dpred, B * T_out * ldd).reshape(B, T_out, ldd)
    D[...] = 0.0
    D[..., :Cc] = np.where(Y != 0, _rn((k * (Pm - Y)).astype(np.float32)), 0.0)      # operand of the proj dgrad / wgrad
    return 0


def _lens_mse(pred, tgt, lens, B, T, log1p_tgt, acc, stream=None):
    Pm, Y = flat(pred, B * T).reshape(B, T), flat(tgt, B * T).reshape(B, T)
    if log1p_tgt:
        Y = np.log(Y + np.float32(1.0))
    m = _lens_mask(lens, B, T)
    A = flat(acc, 2, np.float64)
    A[0] += float((((Pm - Y).astype(np.float32) ** 2).astype(np.float64) * m).sum())
    A[1] += float(m.sum())
    return 0


def _lens_mse_grad(pred, tgt, lens, B, T, log1p_tgt, acc, scale, dpred, stream=None):
    Pm, Y = flat(pred, B * T).reshape(B, T), flat(tgt, B * T).reshape(B, T)
    if log1p_tgt:
        Y = np.log(Y + np.float32(1.0))
    k = np.float32(2.0 * scale / flat(acc, 2, np.float64)[1])
    flat(dpred, B * T).reshape(B, T)[...] = np.where(_lens_mask(lens, B, T), k * (Pm - Y), 0.0)
    return 0


_CHUNK = np.dtype([("start", np.int64), ("len", np.int32), ("tensor", np.int32)])


def _chunks(p, n):
    raw = np.ctypeslib.as_array(C.cast(_addr(p), C.POINTER(C.c_uint8)), shape=(int(n) * _CHUNK.itemsize,))
    return raw.view(_CHUNK)


def _merged(cks):
    """Chunks of one tensor are consecutive: merge them into (start, len, tensor) runs (a few hundred instead of thousands)."""
    runs = []
    for ck in cks:
        a, n, t = int(ck["start"]), int(ck["len"]), int(ck["tensor"])
        if runs and runs[-1][2] == t and runs[-1][0] + runs[-1][1] == a:
            runs[-1][1] += n
        else:
            runs.append([a, n, t])
    return runs


def _grad_sqnorm(g, chunks, n_chunks, out, stream=None):
    s = 0.0
    for a, n, _ in _merged(_chunks(chunks, n_chunks)):
        v = flat(_addr(g) + 4 * a, n)
        s += float(np.dot(v, v))                     # (BLAS sdot: blocked fp32 accumulation, ~1e-7 relative)
    flat(out, 1, np.float64)[0] += s
    return 0


def _lamb_step(p, g, m, v, chunks, n_chunks, norms, gnorm_sq, max_norm, lr_dev, b1, b2, eps, wd, p_tf32, stream=None):
    gs = flat(gnorm_sq, 1, np.float64)
    if gs is not None and not np.isfinite(gs[0]):
        return 0
    coef = np.float32(1.0)
    if gs is not None and max_norm > 0:
        c = np.float32(max_norm) / (np.float32(np.sqrt(gs[0])) + np.float32(1e-6))
        coef = min(c, np.float32(1.0))
    cks = _chunks(chunks, n_chunks)
    runs = _merged(cks)
    n_t = int(cks["tensor"].max()) + 1
    N = flat(norms, 2 * n_t, np.float64)
    b1, b2, eps, wd = (np.float32(x) for x in (b1, b2, eps, wd))
    rs = []
    for a, n, t in runs:
        P, G, M, V = (flat(_addr(q) + 4 * a, n) for q in (p, g, m, v))
        gi = G * coef
        M[...] = b1 * M + (np.float32(1.0) - b1) * gi
        V[...] = b2 * V + (np.float32(1.0) - b2) * gi * gi
        r = M / (np.sqrt(V) + eps) + wd * P
        rs.append(r)
        N[2 * t] += float(np.dot(P, P))
        N[2 * t + 1] += float(np.dot(r, r))
    lr = flat(lr_dev, 1)[0]
    for (a, n, t), r in zip(runs, rs):
        wn = min(np.float32(np.sqrt(N[2 * t])), np.float32(10.0))
        rn = np.float32(np.sqrt(N[2 * t + 1]))
        trust = np.float32(1.0) if (wn == 0 or rn == 0) else wn / rn
        P = flat(_addr(p) + 4 * a, n)
        P[...] = P - lr * trust * r
        if _addr(p_tf32):
            flat(_addr(p_tf32) + 4 * a, n)[...] = _rn(P, "weights")
    return 0


def _attn_fwd(qkv, rs, zs, B, T, lens, scale, drop_p, seed, seed_dev, drop_ld, out, o_rs, o_zs, lse, stream=None):
    """include/xva_b200.h xva_attn_fwd: single head, d_head 64, q | k | v in columns 0..191."""
    X = strided(qkv, (B, T, 192), (zs, rs, 1))
    X = (tf32_trunc(X) if TF32 else X).astype(np.float64)          # tcgen05 kind::tf32 reads its fp32 operands truncated
    q, k, v = X[..., :64], X[..., 64:128], X[..., 128:]
    s_ = np.float32(scale) * np.einsum("bid,bjd->bij", q, k)
    keym = _lens_mask(lens, B, T)[:, None, :]
    s_ = np.where(keym, s_, -np.inf)
    mx = s_.max(axis=2, keepdims=True)
    e = np.exp(s_ - mx)
    den = e.sum(axis=2, keepdims=True)
    P = e / den
    idx = (np.arange(B, dtype=np.uint64)[:, None, None] * np.uint64(T) + np.arange(T, dtype=np.uint64)[None, :, None]) * np.uint64(drop_ld) \
        + np.arange(T, dtype=np.uint64)[None, None, :]
    Pd = P * dropout_scale(seed, seed_dev, idx, drop_p)
    if TF32:
        Pd = tf32_rn(Pd.astype(np.float32)).astype(np.float64)     # P is written to tensor memory rounded: the A operand of P.V
    strided(out, (B, T, 64), (o_zs, o_rs, 1))[...] = _rn(np.einsum("bij,bjd->bid", Pd, v).astype(np.float32))
    flat(lse, B * T).reshape(B, T)[...] = (mx + np.log(den))[..., 0].astype(np.float32)
    return 0


def _attn_bwd(qkv, rs, zs, dout, d_rs, d_zs, lse, dsum, B, T, lens, scale, drop_p, seed, seed_dev, drop_ld, dqkv, g_rs, g_zs,
              stream=None):
    X = strided(qkv, (B, T, 192), (zs, rs, 1))
    X = (tf32_trunc(X) if TF32 else X).astype(np.float64)
    q, k, v = X[..., :64], X[..., 64:128], X[..., 128:]
    dO = strided(dout, (B, T, 64), (d_zs, d_rs, 1))
    dO = (tf32_trunc(dO) if TF32 else dO).astype(np.float64)
    L = flat(lse, B * T).reshape(B, T, 1).astype(np.float64)
    Ds = flat(dsum, B * T).reshape(B, T, 1).astype(np.float64)
    sc = np.float64(np.float32(scale))
    keym = _lens_mask(lens, B, T)[:, None, :]
    P = np.where(keym, np.exp(sc * np.einsum("bid,bjd->bij", q, k) - L), 0.0)
    idx = (np.arange(B, dtype=np.uint64)[:, None, None] * np.uint64(T) + np.arange(T, dtype=np.uint64)[None, :, None]) * np.uint64(drop_ld) \
        + np.arange(T, dtype=np.uint64)[None, None, :]
    dm = dropout_scale(seed, seed_dev, idx, drop_p)
    Pdm = P * dm
    dS = P * (np.einsum("bid,bjd->bij", dO, v) * dm - Ds)
    if TF32:                                                        # P and dS are MMA operands, rounded in tensor memory
        Pdm = tf32_rn(Pdm.astype(np.float32)).astype(np.float64)
        dS = tf32_rn(dS.astype(np.float32)).astype(np.float64)
    dV = np.einsum("bij,bid->bjd", Pdm, dO)
    G = strided(dqkv, (B, T, 192), (g_zs, g_rs, 1))
    G[..., :64] = _rn((sc * np.einsum("bij,bjd->bid", dS, k)).astype(np.float32))
    G[..., 64:128] = _rn((sc * np.einsum("bij,bid->bjd", dS, q)).astype(np.float32))
    G[..., 128:] = _rn(dV.astype(np.float32))
    return 0


def _adamw_step(p, g, m, v, n, lr_dev, b1, b2, eps, wd, step, step_dev, stream=None):
    """csrc/elemwise.cu adamw_kernel: torch.optim.AdamW over a flat arena."""
    P, G, M, V = (flat(q, n) for q in (p, g, m, v))
    lr = flat(lr_dev, 1)[0]
    t = float(flat(step_dev, 1, np.uint64)[0]) if _addr(step_dev) else float(step)
    b1, b2, eps, wd = (np.float32(x) for x in (b1, b2, eps, wd))
    bc1, bc2 = np.float32(1.0 - float(b1) ** t), np.float32(1.0 - float(b2) ** t)
    M[...] = b1 * M + (np.float32(1.0) - b1) * G
    V[...] = b2 * V + (np.float32(1.0) - b2) * G * G
    P[...] = P * (np.float32(1.0) - lr * wd) - (lr / bc1) * (M / (np.sqrt(V) / np.sqrt(bc2) + eps))
    return 0


TABLE.update({
    "xva_adamw_step": _adamw_step,
    "xva_embed_pos": _embed_pos, "xva_embed_bwd": _embed_bwd, "xva_scalar_conv_add": _scalar_conv_add,
    "xva_scalar_conv_bwd": _scalar_conv_bwd, "xva_rowdot_fwd": _rowdot_fwd, "xva_rowdot_bwd": _rowdot_bwd,
    "xva_regulate_len_scan": _regulate_scan, "xva_regulate_len_fwd": _regulate_fwd, "xva_regulate_len_bwd": _regulate_bwd,
    "xva_average_pitch": _average_pitch, "xva_mel_mse": _mel_mse, "xva_mel_mse_grad": _mel_mse_grad, "xva_lens_mse": _lens_mse,
    "xva_lens_mse_grad": _lens_mse_grad, "xva_grad_sqnorm": _grad_sqnorm, "xva_lamb_step": _lamb_step,
    "xva_attn_fwd": _attn_fwd, "xva_attn_bwd": _attn_bwd,
})


# ------------------------------------------------------------------------------------------------ HiFi-GAN generator / WaveNet
def _wn_table(table, n_desc):
    from xva_trainer_b200 import capi
    return (capi.WnDesc * int(n_desc)).from_address(_addr(table))


def _wn_index(d, r, c, j):
    if d.flags & 1:                                    # XVA_WN_TRANSPOSED
        return d.tap_off[j] + c * d.ld + r
    return d.tap_off[j] + r * d.ld + ((r // d.og) % d.f) * d.cg + c


def _wn_pack_fwd(table, n_desc, total_rows, max_inner, stream=None):
    """include/xva_b200.h xva_wn_pack_fwd: w = g * v / ||v|| per output row (w = v for XVA_WN_PLAIN), scattered into the
    packed arena in the tap-GEMM layout."""
    for d in _wn_table(table, n_desc):
        rows, inner, k = d.rows, d.inner, d.k
        c2 = inner // k
        v = flat(d.v, rows * inner).reshape(rows, c2, k).astype(np.float64)
        if d.flags & 4:                                # XVA_WN_PLAIN
            w = v
        else:
            g = flat(d.g, rows).astype(np.float64)
            w = v * (g / np.sqrt((v.reshape(rows, -1) ** 2).sum(axis=1)))[:, None, None]
        r, c, j = np.meshgrid(np.arange(rows), np.arange(c2), np.arange(k), indexing="ij")
        taps = np.array([d.tap_off[t] for t in range(k)], dtype=np.int64)
        if d.flags & 1:
            idx = taps[j] + c * d.ld + r
        else:
            idx = taps[j] + r * d.ld + ((r // d.og) % d.f) * d.cg + c
        dst = flat(d.dst, int(idx.max()) + 1)
        wv = w.reshape(-1).astype(np.float32)
        dst[idx.reshape(-1)] = wv if (d.flags & 2) else _rn(wv, "weights")   # XVA_WN_NO_ROUND: the fp32 first layer
    return 0


def _wn_pack_bwd(table, n_desc, total_rows, max_inner, stream=None):
    for d in _wn_table(table, n_desc):
        rows, inner, k = d.rows, d.inner, d.k
        c2 = inner // k
        v = flat(d.v, rows * inner).reshape(rows, c2, k).astype(np.float64)
        r, c, j = np.meshgrid(np.arange(rows), np.arange(c2), np.arange(k), indexing="ij")
        taps = np.array([d.tap_off[t] for t in range(k)], dtype=np.int64)
        idx = (taps[j] + c * d.ld + r) if (d.flags & 1) else (taps[j] + r * d.ld + ((r // d.og) % d.f) * d.cg + c)
        dW = flat(d.ddst, int(idx.max()) + 1)[idx.reshape(-1)].reshape(rows, c2, k).astype(np.float64)
        dv = flat(d.dv, rows * inner).reshape(rows, c2, k)
        if d.flags & 4:
            dv += dW.astype(np.float32)
            continue
        ss = (v.reshape(rows, -1) ** 2).sum(axis=1)
        dot = (v * dW).reshape(rows, -1).sum(axis=1)
        g = flat(d.g, rows).astype(np.float64)
        scale = g / np.sqrt(ss)
        dv += (scale[:, None, None] * dW - (scale * dot / ss)[:, None, None] * v).astype(np.float32)
        flat(d.dg, rows)[...] += (dot / np.sqrt(ss)).astype(np.float32)
    return 0


def _mean3_lrelu(y0, y1, y2, n, slope, out, stream=None):
    m = ((flat(y0, n) + flat(y1, n)) + flat(y2, n)) * np.float32(1.0 / 3.0)
    flat(out, n)[...] = _rn(np.where(m > 0, m, np.float32(slope) * m).astype(np.float32))
    return 0


def _sum3(a, b, c, n, out, stream=None):
    flat(out, n)[...] = _rn(flat(a, n) + flat(b, n) + flat(c, n))
    return 0


def _tanh_bwd(dy, y, rows, ld, out, stream=None):
    O = flat(out, rows * ld).reshape(rows, ld)
    O[...] = 0.0
    t = flat(y, rows)
    O[:, 0] = _rn(flat(dy, rows) * (np.float32(1.0) - t * t))
    return 0


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _gated_act_fwd(x_in, rows, H, ld_in, acts, stream=None):
    X = strided(x_in, (rows, 2 * H), (ld_in, 1)).astype(np.float64)
    flat(acts, rows * H).reshape(rows, H)[...] = _rn((np.tanh(X[:, :H]) * _sig(X[:, H:])).astype(np.float32))
    return 0


def _gated_act_bwd(dacts, x_in, rows, H, ld_in, dx_in, stream=None):
    X = strided(x_in, (rows, 2 * H), (ld_in, 1)).astype(np.float64)
    d = flat(dacts, rows * H).reshape(rows, H).astype(np.float64)
    t, sg = np.tanh(X[:, :H]), _sig(X[:, H:])
    O = flat(dx_in, rows * 2 * H).reshape(rows, 2 * H)
    O[:, :H] = _rn((d * sg * (1.0 - t * t)).astype(np.float32))
    O[:, H:] = _rn((d * t * sg * (1.0 - sg)).astype(np.float32))
    return 0


def _vits_sample_fwd(stats, eps, lens, B, T, Cc, z, stream=None):
    S = flat(stats, B * T * 2 * Cc).reshape(B, T, 2 * Cc)
    E = flat(eps, B * T * Cc).reshape(B, T, Cc)
    live = _lens_mask(lens, B, T)[:, :, None]
    flat(z, B * T * Cc).reshape(B, T, Cc)[...] = np.where(live, S[..., :Cc] + E * np.exp(S[..., Cc:]), 0.0)
    return 0


def _vits_sample_bwd(dz, eps, stats, lens, B, T, Cc, dstats, stream=None):
    S = flat(stats, B * T * 2 * Cc).reshape(B, T, 2 * Cc)
    E = flat(eps, B * T * Cc).reshape(B, T, Cc)
    D = flat(dz, B * T * Cc).reshape(B, T, Cc)
    live = _lens_mask(lens, B, T)[:, :, None]
    O = flat(dstats, B * T * 2 * Cc).reshape(B, T, 2 * Cc)
    O[..., :Cc] = _rn(np.where(live, D, 0.0).astype(np.float32))
    O[..., Cc:] = _rn(np.where(live, D * E * np.exp(S[..., Cc:]), 0.0).astype(np.float32))
    return 0


TABLE.update({"xva_wn_pack_fwd": _wn_pack_fwd, "xva_wn_pack_bwd": _wn_pack_bwd, "xva_mean3_lrelu": _mean3_lrelu, "xva_sum3": _sum3,
              "xva_tanh_bwd": _tanh_bwd, "xva_gated_act_fwd": _gated_act_fwd, "xva_gated_act_bwd": _gated_act_bwd,
              "xva_vits_sample_fwd": _vits_sample_fwd, "xva_vits_sample_bwd": _vits_sample_bwd})


# ------------------------------------------------------------------------------------------------ discriminators, mel, GAN losses
def _c1_src(z, q, xs_b, xs_q, xs_c, P, Lsrc):
    b, c = z // P, z % P
    idx = q * xs_q + c * xs_c
    idx = np.where(idx >= Lsrc, 2 * (Lsrc - 1) - idx, idx)
    return b * xs_b + idx


def _c1_window(x, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout):
    """xw[z, t, j] = sequence value at q = t * s - pad + j (zero outside [0, L)) and the flat source index of every q."""
    n_src = (Z // P - 1) * xs_b + Lsrc
    X = flat(x, n_src)
    q = np.arange(Lout)[:, None] * s - pad + np.arange(k)[None, :]                   # [Lout, k]
    ok = (q >= 0) & (q < L)
    src = _c1_src(np.arange(Z)[:, None, None], np.clip(q, 0, L - 1)[None], xs_b, xs_q, xs_c, P, Lsrc)
    return np.where(ok[None], X[src], 0.0).astype(np.float64), src, ok, X


def _conv_c1_fwd(x, xs_b, xs_q, xs_c, P, Lsrc, L, w, bias, k, s, pad, Z, Lout, Lout_p, Cout, slope, out, stream=None):
    xw, _, _, _ = _c1_window(x, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout)
    Wt = flat(w, Cout * k).reshape(Cout, k).astype(np.float64)
    v = (xw @ Wt.T + flat(bias, Cout)[None, None, :]).astype(np.float32)
    O = flat(out, Z * Lout_p * Cout).reshape(Z, Lout_p, Cout)
    O[...] = 0.0
    O[:, :Lout] = _rn(np.where(v > 0, v, np.float32(slope) * v).astype(np.float32))      # operand of the next convolution
    return 0


def _conv_c1_bwd_w(dpre, x, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout, dw, db, stream=None):
    xw, _, _, _ = _c1_window(x, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout)
    D = flat(dpre, Z * Lout_p * Cout).reshape(Z, Lout_p, Cout)[:, :Lout].astype(np.float64)
    flat(dw, Cout * k).reshape(Cout, k)[...] += np.einsum("ztc,ztj->cj", D, xw).astype(np.float32)
    flat(db, Cout)[...] += D.sum(axis=(0, 1)).astype(np.float32)
    return 0


def _conv_c1_bwd_x(dpre, w, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout, scale, dx, stream=None):
    D = flat(dpre, Z * Lout_p * Cout).reshape(Z, Lout_p, Cout)[:, :Lout].astype(np.float64)
    Wt = flat(w, Cout * k).reshape(Cout, k).astype(np.float64)
    contrib = np.einsum("ztc,cj->ztj", D, Wt)                                          # gradient wrt sequence value q = t s - pad + j
    q = np.arange(Lout)[:, None] * s - pad + np.arange(k)[None, :]
    ok = (q >= 0) & (q < L)
    src = _c1_src(np.arange(Z)[:, None, None], np.clip(q, 0, L - 1)[None], xs_b, xs_q, xs_c, P, Lsrc)
    n_src = (Z // P - 1) * xs_b + Lsrc
    acc = np.zeros(n_src, np.float64)
    np.add.at(acc, src[np.broadcast_to(ok[None], src.shape)], contrib[np.broadcast_to(ok[None], src.shape)])
    flat(dx, n_src)[...] += (np.float32(scale) * acc).astype(np.float32)
    return 0


def _avgpool4_fwd(x, B, L, out, stream=None):
    Lout = L // 2 + 1
    X = np.pad(flat(x, B * L).reshape(B, L), ((0, 0), (2, 2)))
    i = np.arange(Lout)
    flat(out, B * Lout).reshape(B, Lout)[...] = np.float32(0.25) * (X[:, 2 * i] + X[:, 2 * i + 1] + X[:, 2 * i + 2] + X[:, 2 * i + 3])
    return 0


def _avgpool4_bwd(dout, B, L, dx, stream=None):
    Lout = L // 2 + 1
    D = flat(dout, B * Lout).reshape(B, Lout)
    acc = np.zeros((B, L + 4), np.float32)
    i = np.arange(Lout)
    for d in range(4):
        np.add.at(acc, (slice(None), 2 * i + d), np.float32(0.25) * D)
    flat(dx, B * L).reshape(B, L)[...] = acc[:, 2:L + 2]
    return 0


def _zero_tail_rows(x, Z, Lp, Lvalid, Cc, stream=None):
    flat(x, Z * Lp * Cc).reshape(Z, Lp, Cc)[:, Lvalid:] = 0.0
    return 0


def _sn_table(table, n_desc):
    from xva_trainer_b200 import capi
    return (capi.SnDesc * int(n_desc)).from_address(_addr(table))


def _sn_idx(d):
    rows, inner, k = d.rows, d.inner, d.k
    c2 = inner // k
    r, c, j = np.meshgrid(np.arange(rows), np.arange(c2), np.arange(k), indexing="ij")
    taps = np.array([d.tap_off[t] for t in range(k)], dtype=np.int64)
    return (taps[j] + c * d.ld + r) if (d.flags & 1) else (taps[j] + r * d.ld + ((r // d.og) % d.f) * d.cg + c)


def _sn_pack_fwd(table, n_desc, total_rows, total_blocks, max_inner, training, stream=None):
    """include/xva_b200.h xva_sn_pack_fwd: one power iteration (training), sigma = u . (W v), W / sigma packed."""
    for d in _sn_table(table, n_desc):
        rows, inner, k = d.rows, d.inner, d.k
        Wm = flat(d.w, rows * inner).reshape(rows, inner).astype(np.float64)
        u, v = flat(d.u, rows), flat(d.v, inner)
        if training:
            t = Wm.T @ u.astype(np.float64)
            vn = (t / max(np.sqrt((t * t).sum()), 1e-12)).astype(np.float32)
            v[...] = vn
            s_ = Wm @ vn.astype(np.float64)
            un = (s_ / max(np.sqrt((s_ * s_).sum()), 1e-12)).astype(np.float32)
            u[...] = un
            sigma = float((un.astype(np.float64) * s_).sum())
        else:
            s_ = Wm @ v.astype(np.float64)
            sigma = float((u.astype(np.float64) * s_).sum())
        flat(d.u_sav, rows)[...] = u
        flat(d.v_sav, inner)[...] = v
        chunks = (rows + 63) // 64
        flat(d.work, chunks * inner + rows + 2)[chunks * inner + rows] = np.float32(sigma)
        idx = _sn_idx(d)
        dst = flat(d.dst, int(idx.max()) + 1)
        wv = (Wm.reshape(rows, inner // k, k) / np.float32(sigma)).reshape(-1).astype(np.float32)
        dst[idx.reshape(-1)] = wv if (d.flags & 2) else _rn(wv, "weights")
    return 0


def _sn_pack_bwd(table, n_desc, total_rows, total_blocks, max_inner, stream=None):
    for d in _sn_table(table, n_desc):
        rows, inner, k = d.rows, d.inner, d.k
        Wm = flat(d.w, rows * inner).reshape(rows, inner).astype(np.float64)
        chunks = (rows + 63) // 64
        sigma = float(flat(d.work, chunks * inner + rows + 2)[chunks * inner + rows])
        idx = _sn_idx(d)
        dW = flat(d.ddst, int(idx.max()) + 1)[idx.reshape(-1)].reshape(rows, inner).astype(np.float64)
        coef = float((dW * Wm).sum()) / (sigma * sigma)
        u, v = flat(d.u_sav, rows).astype(np.float64), flat(d.v_sav, inner).astype(np.float64)
        flat(d.dw, rows * inner).reshape(rows, inner)[...] += (dW / sigma - coef * np.outer(u, v)).astype(np.float32)
    return 0


def _reflect_index(t, n):
    t = np.where(t < 0, -t, t)
    return np.where(t >= n, 2 * (n - 1) - t, t)


def _reflect_pad_fwd(y, B, n, pad, out, stream=None):
    Y = flat(y, B * n).reshape(B, n)
    flat(out, B * (n + 2 * pad)).reshape(B, n + 2 * pad)[...] = _rn(Y[:, _reflect_index(np.arange(n + 2 * pad) - pad, n)])
    return 0


def _reflect_pad_bwd(dyp, B, n, pad, dy, stream=None):
    D = flat(dyp, B * (n + 2 * pad)).reshape(B, n + 2 * pad)
    acc = np.zeros((B, n), np.float32)
    np.add.at(acc, (slice(None), _reflect_index(np.arange(n + 2 * pad) - pad, n)), D)
    flat(dy, B * n).reshape(B, n)[...] = acc
    return 0


def _spec_mag_fwd(spec, rows, nb, ld_s, ld_m, eps, mag, stream=None):
    S = flat(spec, rows * ld_s).reshape(rows, ld_s)
    p = S[:, :nb] ** 2 + S[:, nb:2 * nb] ** 2
    M = flat(mag, rows * ld_m).reshape(rows, ld_m)
    M[...] = 0.0
    M[:, :nb] = _rn(np.sqrt(p + np.float32(eps) if eps >= 0 else np.maximum(p, np.float32(-eps))).astype(np.float32))
    return 0


def _spec_mag_bwd(dmag, spec, rows, nb, ld_s, ld_m, eps, dspec, stream=None):
    S = flat(spec, rows * ld_s).reshape(rows, ld_s)
    D = flat(dmag, rows * ld_m).reshape(rows, ld_m)[:, :nb]
    p = S[:, :nb] ** 2 + S[:, nb:2 * nb] ** 2
    if eps >= 0:
        m, live = np.sqrt(p + np.float32(eps)), np.ones_like(p, bool)
    else:
        m, live = np.sqrt(np.maximum(p, 1e-30)), p >= np.float32(-eps)
    O = flat(dspec, rows * ld_s).reshape(rows, ld_s)
    O[...] = 0.0
    O[:, :nb] = _rn(np.where(live, D * S[:, :nb] / m, 0.0).astype(np.float32))
    O[:, nb:2 * nb] = _rn(np.where(live, D * S[:, nb:2 * nb] / m, 0.0).astype(np.float32))
    return 0


def _log_clamp_fwd(x, n, lo, out, stream=None):
    flat(out, n)[...] = np.log(np.maximum(flat(x, n), np.float32(lo)))
    return 0


def _log_clamp_bwd(dy, x, n, lo, dx, stream=None):
    X = flat(x, n)
    flat(dx, n)[...] = _rn(np.where(X >= np.float32(lo), flat(dy, n) / np.where(X >= np.float32(lo), X, 1.0), 0.0).astype(np.float32))
    return 0


def _reduce_loss(a, b, n, kind, c, acc, stream=None):
    A = flat(a, n)
    if kind == 0:
        flat(acc, 1, np.float64)[0] += float(np.abs(A - flat(b, n)).astype(np.float64).sum())
    else:
        dd = (np.float32(c) - A).astype(np.float32)
        flat(acc, 1, np.float64)[0] += float((dd * dd).astype(np.float64).sum())
    return 0


def _l1_grad_values(a, b, n, scale, gate_slope):
    A, Bv = flat(a, n), flat(b, n)
    dd = Bv - A
    g_ = np.where(dd > 0, np.float32(scale), np.where(dd < 0, np.float32(-scale), np.float32(0.0)))
    return np.where(Bv > 0, g_, g_ * np.float32(gate_slope)).astype(np.float32), dd


def _loss_grad(a, b, n, kind, c, scale, gate_slope, accumulate, out, stream=None):
    if kind == 0:
        g_, _ = _l1_grad_values(a, b, n, scale, gate_slope)
    else:
        g_ = (np.float32(scale) * np.float32(2.0) * (flat(a, n) - np.float32(c))).astype(np.float32)
    O = flat(out, n)
    O[...] = O + g_ if accumulate else g_
    return 0


def _l1_loss_grad(a, b, n, scale, gate_slope, acc, out, stream=None):
    g_, dd = _l1_grad_values(a, b, n, scale, gate_slope)
    flat(acc, 1, np.float64)[0] += float(np.abs(dd).astype(np.float64).sum())
    flat(out, n)[...] = g_
    return 0


TABLE.update({"xva_conv_c1_fwd": _conv_c1_fwd, "xva_conv_c1_bwd_w": _conv_c1_bwd_w, "xva_conv_c1_bwd_x": _conv_c1_bwd_x,
              "xva_avgpool4_fwd": _avgpool4_fwd, "xva_avgpool4_bwd": _avgpool4_bwd, "xva_zero_tail_rows": _zero_tail_rows,
              "xva_sn_pack_fwd": _sn_pack_fwd, "xva_sn_pack_bwd": _sn_pack_bwd, "xva_reflect_pad_fwd": _reflect_pad_fwd,
              "xva_reflect_pad_bwd": _reflect_pad_bwd, "xva_spec_mag_fwd": _spec_mag_fwd, "xva_spec_mag_bwd": _spec_mag_bwd,
              "xva_log_clamp_fwd": _log_clamp_fwd, "xva_log_clamp_bwd": _log_clamp_bwd, "xva_reduce_loss": _reduce_loss,
              "xva_loss_grad": _loss_grad, "xva_l1_loss_grad": _l1_loss_grad})


# ------------------------------------------------------------------------------------------------ stage-1 aligner, xVAPitch alignment
def _attn_score_fwd(q, ldq, k, ldk, prior, in_lens, B, Tm, Tt, Cc, logprob, soft, stream=None):
    Q = strided(q, (B, Tm, Cc), (Tm * ldq, ldq, 1)).astype(np.float64)
    Kk = strided(k, (B, Tt, Cc), (Tt * ldk, ldk, 1)).astype(np.float64)
    Pr = flat(prior, B * Tm * Tt).reshape(B, Tm, Tt).astype(np.float64)
    D = -0.0005 * ((Q[:, :, None, :] - Kk[:, None, :, :]) ** 2).sum(-1)
    lse = np.log(np.exp(D - D.max(2, keepdims=True)).sum(2, keepdims=True)) + D.max(2, keepdims=True)
    lp = (D - lse) + np.log(Pr + 1e-8)
    flat(logprob, B * Tm * Tt).reshape(B, Tm, Tt)[...] = lp.astype(np.float32)
    nk = np.clip(flat(in_lens, B, np.int32), 0, Tt)
    live = np.arange(Tt)[None, None, :] < nk[:, None, None]
    m = np.where(live, lp, -np.inf)
    e = np.exp(m - m.max(2, keepdims=True)) * live
    flat(soft, B * Tm * Tt).reshape(B, Tm, Tt)[...] = (e / e.sum(2, keepdims=True)).astype(np.float32)
    return 0


def _attn_score_bwd(g, logprob, prior, q, ldq, k, ldk, B, Tm, Tt, Cc, dD, dq, lddq, dk, lddk, stream=None):
    Gm = flat(g, B * Tm * Tt).reshape(B, Tm, Tt).astype(np.float64)
    sm = np.exp(flat(logprob, B * Tm * Tt).reshape(B, Tm, Tt).astype(np.float64)
                - np.log(flat(prior, B * Tm * Tt).reshape(B, Tm, Tt).astype(np.float64) + 1e-8))
    d = Gm - sm * Gm.sum(2, keepdims=True)
    Q = strided(q, (B, Tm, Cc), (Tm * ldq, ldq, 1)).astype(np.float64)
    Kk = strided(k, (B, Tt, Cc), (Tt * ldk, ldk, 1)).astype(np.float64)
    flat(dD, B * Tm * Tt).reshape(B, Tm, Tt)[...] = d.astype(np.float32)            # (may alias g: written after it was read)
    diff = Q[:, :, None, :] - Kk[:, None, :, :]                                      # q - k
    strided(dq, (B, Tm, Cc), (Tm * lddq, lddq, 1))[...] = (-0.001 * np.einsum("btj,btjc->btc", d, diff)).astype(np.float32)
    strided(dk, (B, Tt, Cc), (Tt * lddk, lddk, 1))[...] = (0.001 * np.einsum("btj,btjc->bjc", d, diff)).astype(np.float32)
    return 0


def _attn_ctc(logprob, in_lens, out_lens, B, Tm, Tt, blank_logprob, workspace, workspace_bytes, cost, grad, stream=None):
    """AttentionCTCLoss (fastpitch/attn_loss_function.py:20-44) per utterance through torch's own CTC and autograd -- an
    implementation independent of the kernel's fp64 recursion."""
    import torch
    import torch.nn.functional as Fn
    LP = torch.from_numpy(flat(logprob, B * Tm * Tt).reshape(B, Tm, Tt).copy()).double().requires_grad_(True)
    il, ol = flat(in_lens, B, np.int32), flat(out_lens, B, np.int32)
    total = 0.0
    costs = np.zeros(B)
    for b in range(B):
        L, T = int(il[b]), int(ol[b])
        rows = torch.cat([torch.full((T, 1), float(blank_logprob), dtype=torch.float64), LP[b, :T, :L]], 1)
        lsm = torch.log_softmax(rows, dim=1)[:, None, :]
        c = Fn.ctc_loss(lsm, torch.arange(1, L + 1)[None], torch.tensor([T]), torch.tensor([L]), blank=0, reduction="mean",
                        zero_infinity=True)
        costs[b] = float(c.detach())
        total = total + c / B
    total.backward()
    flat(cost, B, np.float64)[...] = costs
    flat(grad, B * Tm * Tt).reshape(B, Tm, Tt)[...] = LP.grad.float().numpy()
    return 0


def _attn_bin_loss(hard, soft, rows, Tt, eps, acc, stream=None):
    Hd, Sf = flat(hard, rows * Tt), flat(soft, rows * Tt)
    A = flat(acc, 2, np.float64)
    A[0] += float(np.log(np.maximum(Sf[Hd == 1], np.float32(eps)).astype(np.float64)).sum())
    A[1] += float(Hd.astype(np.float64).sum())
    return 0


def _attn_grad_combine(gctc, hard, soft, acc, a, bw, eps, rows, Tt, g, stream=None):
    Gc = flat(gctc, rows * Tt).reshape(rows, Tt)
    out = np.float32(a) * Gc
    if _addr(hard) and bw != 0.0:
        Hd, Sf = flat(hard, rows * Tt).reshape(rows, Tt), flat(soft, rows * Tt).reshape(rows, Tt)
        hp = Hd * (Sf >= np.float32(eps))
        out = out + np.float32(bw / flat(acc, 2, np.float64)[1]) * (Sf * hp.sum(1, keepdims=True) - hp)
    flat(g, rows * Tt).reshape(rows, Tt)[...] = out.astype(np.float32)
    return 0


def _mas_log(attn, n, out, stream=None):
    with np.errstate(divide="ignore"):
        flat(out, n)[...] = np.log(flat(attn, n).astype(np.float64)).astype(np.float32)
    return 0


def _mas_width1(attn, in_lens, out_lens, B, Tm, Tt, is_log, hard, durs, stream=None):
    """The integer path is NOT restated a second time: it is the oracle's own search (oracle.fastpitch.b_mas, pinned to the
    reference's numba b_mas; oracle.vits.maximum_path, pinned to xVAPitch's) -- on the CPU these calls only check the host
    code around the kernel; the kernel itself is compared with the same functions bit for bit in tests/test_mas_gpu.py."""
    import torch
    from oracle import fastpitch as ofp, vits as ov
    A = flat(attn, B * Tm * Tt).reshape(B, Tm, Tt)
    il, ol = flat(in_lens, B, np.int32), flat(out_lens, B, np.int32)
    if is_log & 2:
        path = ov.maximum_path(torch.from_numpy(np.ascontiguousarray(A.transpose(0, 2, 1))), il, ol).numpy().transpose(0, 2, 1)
    else:
        path = ofp.b_mas(A[:, None].copy(), il, ol, is_log=bool(is_log & 1))[:, 0]
    flat(hard, B * Tm * Tt).reshape(B, Tm, Tt)[...] = path
    flat(durs, B * Tt, np.int32).reshape(B, Tt)[...] = path.sum(1).astype(np.int32)
    return 0


def _vits_logp_operands(m_p, logs_p, z_p, B, Tt, Ts, Cc, tok, frm, stream=None):
    K = 2 * Cc + 32
    M, Lg = (flat(t, B * Tt * Cc).reshape(B, Tt, Cc).astype(np.float64) for t in (m_p, logs_p))
    Zp = flat(z_p, B * Ts * Cc).reshape(B, Ts, Cc).astype(np.float64)
    o = np.exp(-2.0 * Lg)
    Tk = flat(tok, B * Tt * K).reshape(B, Tt, K)
    Tk[...] = 0.0
    Tk[..., :Cc] = o
    Tk[..., Cc:2 * Cc] = M * o
    Tk[..., 2 * Cc] = (-0.5 * np.log(2 * np.pi) - Lg - 0.5 * M * M * o).sum(-1)
    Fr = flat(frm, B * Ts * K).reshape(B, Ts, K)
    Fr[...] = 0.0
    Fr[..., :Cc] = -0.5 * Zp * Zp
    Fr[..., Cc:2 * Cc] = Zp
    Fr[..., 2 * Cc] = 1.0
    return 0


def _vits_kl(z_p, logs_q, m_p, logs_p, lens, B, T, Cc, scale, acc, dz_p, dlogs_q, dm_p, dlogs_p, stream=None):
    Z, Lq, M, Lp = (flat(t, B * T * Cc).reshape(B, T, Cc).astype(np.float64) for t in (z_p, logs_q, m_p, logs_p))
    live = _lens_mask(lens, B, T)[:, :, None]
    e = np.exp(-2.0 * Lp)
    d = Z - M
    flat(acc, 1, np.float64)[0] += float(((Lp - Lq - 0.5 + 0.5 * d * d * e) * live).sum())
    k = float(scale) / float(flat(lens, B, np.int32).sum())
    for ptr, val in ((dz_p, d * e), (dlogs_q, -np.ones_like(d)), (dm_p, -d * e), (dlogs_p, 1.0 - d * d * e)):
        flat(ptr, B * T * Cc).reshape(B, T, Cc)[...] = (k * val * live).astype(np.float32)
    return 0


TABLE.update({"xva_attn_score_fwd": _attn_score_fwd, "xva_attn_score_bwd": _attn_score_bwd, "xva_attn_ctc": _attn_ctc,
              "xva_attn_bin_loss": _attn_bin_loss, "xva_attn_grad_combine": _attn_grad_combine, "xva_mas_log": _mas_log,
              "xva_mas_width1": _mas_width1, "xva_vits_logp_operands": _vits_logp_operands, "xva_vits_kl": _vits_kl})
