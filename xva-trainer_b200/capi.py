"""ctypes binding of libxva_b200.so (C ABI declared in include/xva_b200.h).

This is the only place that touches the shared library. Every function raises XvaError on a non-zero status;
there is no fallback implementation: if the library is missing or the device is not sm_100 the call fails.
"""
import ctypes as C
import os

from . import build as _build

XVA_MAX_TAPS = 48

GEMM_RELU = 1 << 0
GEMM_LN = 1 << 1
GEMM_DROP_PRE = 1 << 2
GEMM_DROP_POST = 1 << 3
GEMM_ATOMIC = 1 << 4
GEMM_LRELU_GATE = 1 << 5
GEMM_ROUND_OUT = 1 << 6
GEMM_TANH = 1 << 7
GEMM_SOFTMAX_BWD = 1 << 8
GEMM_HALO = 1 << 9

c_float_p = C.c_void_p  # device pointers travel as integers


class XvaError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of struct xva_gemm_args (include/xva_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32), ("Z", C.c_int32), ("R", C.c_int32), ("M", C.c_int32), ("N", C.c_int32),
        ("K", C.c_int32), ("taps", C.c_int32), ("shift", C.c_int32 * XVA_MAX_TAPS), ("ZR", C.c_int32),
        ("split", C.c_int32),
        ("a", C.c_void_p), ("a_rs", C.c_int64), ("a_zs", C.c_int64), ("a_rows", C.c_int32), ("_pad0", C.c_int32),
        ("b", C.c_void_p), ("b_rs", C.c_int64), ("b_zs", C.c_int64), ("b_rows", C.c_int32), ("b_nz", C.c_int32),
        ("b_tap_z", C.c_int32), ("b_batch_z", C.c_int32),
        ("out", C.c_void_p), ("o_rs", C.c_int64), ("o_zs", C.c_int64), ("o_js", C.c_int64),
        ("alpha", C.c_float), ("flags", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("r_rs", C.c_int64), ("r_zs", C.c_int64),
        ("gate", C.c_void_p), ("g_rs", C.c_int64), ("g_zs", C.c_int64), ("gate_slope", C.c_float),
        ("act_slope", C.c_float),
        ("lens", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("ln_eps", C.c_float),
        ("_pad2", C.c_int32),
        ("out_pre", C.c_void_p), ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p), ("drop_p", C.c_float),
        ("out_act_slope", C.c_float), ("seed", C.c_uint64), ("out_act", C.c_void_p),
        ("a_col", C.c_int32 * XVA_MAX_TAPS), ("seed_dev", C.c_void_p),
        ("groups", C.c_int32), ("grp_step", C.c_int32),
        ("rowvec", C.c_void_p), ("drop_ld", C.c_int32), ("_pad3", C.c_int32),
    ]


class WnDesc(C.Structure):
    """Mirror of struct xva_wn_desc (include/xva_b200.h)."""
    _fields_ = [
        ("v", C.c_void_p), ("g", C.c_void_p), ("dv", C.c_void_p), ("dg", C.c_void_p), ("dst", C.c_void_p),
        ("ddst", C.c_void_p),
        ("rows", C.c_int32), ("inner", C.c_int32), ("k", C.c_int32), ("flags", C.c_int32),
        ("ld", C.c_int32), ("og", C.c_int32), ("f", C.c_int32), ("cg", C.c_int32),
        ("row_start", C.c_int32), ("_pad", C.c_int32),
        ("tap_off", C.c_int64 * XVA_MAX_TAPS),
    ]


class SnDesc(C.Structure):
    """Mirror of struct xva_sn_desc (include/xva_b200.h)."""
    _fields_ = [
        ("w", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("u_sav", C.c_void_p), ("v_sav", C.c_void_p),
        ("dw", C.c_void_p), ("dst", C.c_void_p), ("ddst", C.c_void_p), ("work", C.c_void_p),
        ("rows", C.c_int32), ("inner", C.c_int32), ("k", C.c_int32), ("flags", C.c_int32),
        ("ld", C.c_int32), ("og", C.c_int32), ("f", C.c_int32), ("cg", C.c_int32),
        ("row_start", C.c_int32), ("blk_start", C.c_int32),
        ("tap_off", C.c_int64 * XVA_MAX_TAPS),
    ]


WN_TRANSPOSED, WN_NO_ROUND, WN_PLAIN = 1, 2, 4

# name -> (restype, argtypes); must list every symbol include/xva_b200.h declares (tests/test_abi.py checks it)
_I, _F, _P, _U64, _I64 = C.c_int, C.c_float, C.c_void_p, C.c_uint64, C.c_int64
PROTOTYPES = {
    "xva_abi_version": (_I, []),
    "xva_last_error": (C.c_char_p, []),
    "xva_device_check": (_I, [_I]),
    "xva_sizeof_gemm_args": (_I, []),
    "xva_gemm": (_I, [C.POINTER(GemmArgs), _P]),
    "xva_gemm_ref": (_I, [C.POINTER(GemmArgs), _P]),
    "xva_gemm_debug_counters": (_I, [C.POINTER(C.c_longlong * 8)]),
    "xva_regulate_len_scan": (_I, [_P, _I, _I, _F, _I, _P, _P, _P]),
    "xva_regulate_len_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "xva_regulate_len_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "xva_average_pitch": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "xva_rowdot2": (_I, [_P, _P, _I64, _I, _I64, _I64, _P, _P]),
    "xva_attn_fwd": (_I, [_P, _I64, _I64, _I, _I, _P, _F, _F, _U64, _P, _I, _P, _I64, _I64, _P, _P]),
    "xva_attn_bwd": (_I, [_P, _I64, _I64, _P, _I64, _I64, _P, _P, _I, _I, _P, _F, _F, _U64, _P, _I, _P, _I64, _I64, _P]),
    "xva_mas_width1": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "xva_attn_score_fwd": (_I, [_P, _I64, _P, _I64, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "xva_attn_score_bwd": (_I, [_P, _P, _P, _P, _I64, _P, _I64, _I, _I, _I, _I, _P, _P, _I64, _P, _I64, _P]),
    "xva_attn_ctc_workspace_bytes": (_I64, [_I, _I, _I]),
    "xva_attn_ctc": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _I64, _P, _P, _P]),
    "xva_attn_bin_loss": (_I, [_P, _P, _I64, _I, _F, _P, _P]),
    "xva_attn_grad_combine": (_I, [_P, _P, _P, _P, _F, _F, _F, _I64, _I, _P, _P]),
    "xva_mas_log": (_I, [_P, _I64, _P, _P]),
    "xva_softmax_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _F, _U64, _P, _P]),
    "xva_softmax_bwd": (_I, [_P, _P, _I, _I, _I, _I, _F, _F, _U64, _P, _P]),
    "xva_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _F, _U64, _F, _U64, _P, _I, _P]),
    "xva_layernorm_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "xva_counter_add": (_I, [_P, _U64, _P]),
    "xva_colsum": (_I, [_P, _I64, _I, _I64, _P, _P]),
    "xva_embed_pos": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "xva_embed_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "xva_scalar_conv_add": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "xva_scalar_conv_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "xva_rowdot_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "xva_rowdot_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "xva_mel_mse": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "xva_mel_mse_grad": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _F, _P, _P]),
    "xva_lens_mse": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "xva_lens_mse_grad": (_I, [_P, _P, _P, _I, _I, _I, _P, _F, _P, _P]),
    "xva_grad_sqnorm": (_I, [_P, _P, _I, _P, _P]),
    "xva_lamb_step": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _F, _P, _F, _F, _F, _F, _P, _P]),
    "xva_round_tf32": (_I, [_P, _P, _I64, _P]),
    "xva_set_operand_rounding": (_I, [_I]),
    "xva_reflect_pad_fwd": (_I, [_P, _I, _I64, _I, _P, _P]),
    "xva_reflect_pad_bwd": (_I, [_P, _I, _I64, _I, _P, _P]),
    "xva_spec_mag_fwd": (_I, [_P, _I64, _I, _I, _I, _F, _P, _P]),
    "xva_spec_mag_bwd": (_I, [_P, _P, _I64, _I, _I, _I, _F, _P, _P]),
    "xva_log_clamp_fwd": (_I, [_P, _I64, _F, _P, _P]),
    "xva_log_clamp_bwd": (_I, [_P, _P, _I64, _F, _P, _P]),
    "xva_reduce_loss": (_I, [_P, _P, _I64, _I, _F, _P, _P]),
    "xva_loss_grad": (_I, [_P, _P, _I64, _I, _F, _F, _F, _I, _P, _P]),
    "xva_conv_c1_fwd": (_I, [_P, _I64, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "xva_conv_c1_bwd_w": (_I, [_P, _P, _I64, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "xva_conv_c1_bwd_x": (_I, [_P, _P, _I64, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "xva_avgpool4_fwd": (_I, [_P, _I, _I, _P, _P]),
    "xva_avgpool4_bwd": (_I, [_P, _I, _I, _P, _P]),
    "xva_zero_tail_rows": (_I, [_P, _I, _I, _I, _I, _P]),
    "xva_mean3_lrelu": (_I, [_P, _P, _P, _I64, _F, _P, _P]),
    "xva_sum3": (_I, [_P, _P, _P, _I64, _P, _P]),
    "xva_tanh_bwd": (_I, [_P, _P, _I64, _I, _P, _P]),
    "xva_gated_act_fwd": (_I, [_P, _I64, _I, _I64, _P, _P]),
    "xva_gated_act_bwd": (_I, [_P, _P, _I64, _I, _I64, _P, _P]),
    "xva_colsum_items": (_I, [_P, _I, _I, _I, _I64, _I64, _P, _I64, _P]),
    "xva_vits_logp_operands": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "xva_vits_kl": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "xva_vits_sample_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "xva_vits_sample_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "xva_text_embed_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    "xva_text_embed_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P]),
    "xva_rel_band_add": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "xva_rel_band_gather": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "xva_pad_cols": (_I, [_P, _I64, _I, _I, _P, _P]),
    "xva_l1_loss_grad": (_I, [_P, _P, _I64, _F, _F, _P, _P, _P]),
    "xva_sizeof_wn_desc": (_I, []),
    "xva_sizeof_sn_desc": (_I, []),
    "xva_sn_pack_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P]),
    "xva_sn_pack_bwd": (_I, [_P, _I, _I, _I, _I, _P]),
    "xva_wn_pack_fwd": (_I, [_P, _I, _I, _I, _P]),
    "xva_wn_pack_bwd": (_I, [_P, _I, _I, _I, _P]),
    "xva_adamw_step": (_I, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _I, _P, _P]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (once) and return the ctypes handle. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise XvaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                       "(there is no CPU / eager fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.xva_sizeof_sn_desc() != C.sizeof(SnDesc):
        raise XvaError(f"xva_sn_desc layout mismatch: C {lib.xva_sizeof_sn_desc()} vs ctypes {C.sizeof(SnDesc)}")
    if lib.xva_sizeof_wn_desc() != C.sizeof(WnDesc):
        raise XvaError(f"xva_wn_desc layout mismatch: C {lib.xva_sizeof_wn_desc()} vs ctypes {C.sizeof(WnDesc)}")
    if lib.xva_sizeof_gemm_args() != C.sizeof(GemmArgs):
        raise XvaError(f"xva_gemm_args layout mismatch: C {lib.xva_sizeof_gemm_args()} vs ctypes {C.sizeof(GemmArgs)}")
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().xva_last_error()
        raise XvaError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


# kernels enqueued per successful call (everything not listed launches exactly one)
_LAUNCHES = {"xva_lamb_step": 2, "xva_attn_bwd": 2, "xva_attn_score_bwd": 2, "xva_attn_ctc": 3, "xva_gemm_debug_counters": 0, "xva_set_operand_rounding": 0, "xva_abi_version": 0, "xva_last_error": 0, "xva_device_check": 0,
             "xva_sizeof_gemm_args": 0, "xva_sizeof_wn_desc": 0, "xva_sizeof_sn_desc": 0, "xva_sn_pack_fwd": 5, "xva_sn_pack_bwd": 3, "xva_vits_logp_operands": 2}
_launch_count = 0


def reset_launch_count():
    global _launch_count
    _launch_count = 0


def launch_count():
    """CUDA kernels of this library launched through call() since the last reset (bench.py's gpu_launches)."""
    return _launch_count


def call(name, *args):
    global _launch_count
    check(getattr(load(), name)(*args), name)
    _launch_count += _LAUNCHES.get(name, 1)
