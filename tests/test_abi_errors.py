"""Error behaviour of the C ABI (include/xva_b200.h conventions): an unsupported shape or a bad argument is refused on the
host with a negative status and a message in xva_last_error(), before anything is enqueued -- no exception crosses the
boundary, nothing is launched, no CPU fallback takes over. Runs without a GPU: the pointers handed in are never
dereferenced on the host (every check precedes the first CUDA call of its entry point)."""
import ctypes as C

import pytest

ERR_ARG = -1


@pytest.fixture(scope="module")
def abi(lib):
    from xva_trainer_b200 import capi

    buf = (C.c_double * 64)()                      # any non-null, 8-byte aligned address

    def call(name, *args):
        rc = getattr(lib, name)(*args)
        return rc, (lib.xva_last_error() or b"").decode()

    return capi, C.cast(buf, C.c_void_p), call


def test_attn_score_limits(abi):
    capi, p, call = abi
    ok_shape = dict(B=2, Tm=50, Tt=14, Cc=80)
    rc, msg = call("xva_attn_score_fwd", p, 96, p, 96, p, p, 2, 50, 600, 80, p, p, None)
    assert rc == ERR_ARG and "Tt=600" in msg and "512" in msg                      # text positions kept in registers
    rc, msg = call("xva_attn_score_fwd", p, 128, p, 128, p, p, 2, 50, 14, 128, p, p, None)
    assert rc == ERR_ARG and "C=128" in msg
    rc, msg = call("xva_attn_score_fwd", p, 64, p, 96, p, p, 2, 50, 14, 80, p, p, None)
    assert rc == ERR_ARG and "ldq=64" in msg                                       # row pitch shorter than the row
    rc, msg = call("xva_attn_score_fwd", None, 96, p, 96, p, p, 2, 50, 14, 80, p, p, None)
    assert rc == ERR_ARG and "null" in msg
    rc, msg = call("xva_attn_score_fwd", p, 96, p, 96, p, p, 0, 50, 14, 80, p, p, None)
    assert rc == ERR_ARG and "B=0" in msg
    rc, msg = call("xva_attn_score_bwd", p, p, p, p, 96, p, 96, 2, 50, 14, 80, p, p, 64, p, 96, None)
    assert rc == ERR_ARG and "lddq=64" in msg
    assert ok_shape  # (the same calls with legal shapes are the GPU tests of tests/test_stage1_gpu.py)


def test_attn_ctc_workspace_contract(abi):
    capi, p, call = abi
    lib = capi.load()
    need = lib.xva_attn_ctc_workspace_bytes(32, 880, 160)
    assert need == 2 * 32 * 880 * 321 * 8 + 32 * 8 + 32 * 880 * 4                  # alpha + beta tables, nll, row normalisers
    rc, msg = call("xva_attn_ctc", p, p, p, 32, 880, 160, C.c_float(-1.0), p, need - 8, p, p, None)
    assert rc == ERR_ARG and "workspace" in msg and str(need) in msg
    rc, msg = call("xva_attn_ctc", p, p, p, 32, 880, 160, C.c_float(-1.0), C.c_void_p(p.value + 4), need, p, p, None)
    assert rc == ERR_ARG and "aligned" in msg
    rc, msg = call("xva_attn_ctc", p, p, p, 2, 100, 1024, C.c_float(-1.0), p, 1 << 40, p, p, None)
    assert rc == ERR_ARG and "Tt=1024" in msg
    rc, msg = call("xva_attn_ctc", p, p, p, 2, 100, 20, C.c_float(-1.0), None, 1 << 30, p, p, None)
    assert rc == ERR_ARG and "null" in msg


def test_mas_and_binarization_arguments(abi):
    capi, p, call = abi
    rc, msg = call("xva_mas_width1", p, p, p, 2, 20000, 160, 1, p, p, None)
    assert rc == ERR_ARG and "shared memory" in msg                                # Tm x Tt/32 choice bits must fit one SM
    rc, msg = call("xva_mas_width1", p, p, p, 0, 10, 10, 1, p, p, None)
    assert rc == ERR_ARG
    rc, msg = call("xva_mas_log", None, 10, p, None)
    assert rc == ERR_ARG
    rc, msg = call("xva_attn_bin_loss", p, p, 0, 14, C.c_float(1e-12), p, None)
    assert rc == ERR_ARG and "rows=0" in msg
    rc, msg = call("xva_attn_grad_combine", p, p, None, None, C.c_float(1.0), C.c_float(0.5), C.c_float(1e-12), 10, 14, p, None)
    assert rc == ERR_ARG and "hard without soft" in msg


def test_gemm_argument_checks(abi):
    capi, p, call = abi

    def args(**kw):
        g = capi.GemmArgs()
        g.mode, g.taps, g.Z, g.R, g.N, g.K = 0, 1, 2, 64, 64, 64
        g.ZR = g.split = g.b_nz = 1
        g.alpha = 1.0
        g.a = g.b = g.out = p
        g.a_rs = g.b_rs = g.o_rs = 64
        g.a_zs = g.o_zs = 64 * 64
        g.b_zs = 64 * 64
        for k, v in kw.items():
            setattr(g, k, v)
        return g

    rc, msg = call("xva_gemm", C.byref(args(mode=3)), None)
    assert rc == ERR_ARG and "bad mode 3" in msg
    rc, msg = call("xva_gemm", C.byref(args(taps=0)), None)
    assert rc == ERR_ARG and "taps 0" in msg
    rc, msg = call("xva_gemm", C.byref(args(taps=capi.XVA_MAX_TAPS + 1)), None)
    assert rc == ERR_ARG and "taps" in msg
    rc, msg = call("xva_gemm", C.byref(args(R=0)), None)
    assert rc == ERR_ARG and "empty problem" in msg
    g = args()
    g.out = None
    rc, msg = call("xva_gemm", C.byref(g), None)
    assert rc == ERR_ARG and "null operand" in msg
    rc, msg = call("xva_gemm", None, None)
    assert rc == ERR_ARG and "null args" in msg
    # weight gradient with an 80-column MN-major operand in rows of 80 floats: refused (rows of >= 96 are required)
    rc, msg = call("xva_gemm", C.byref(args(mode=2, M=80, N=64, a_rs=80, ZR=2, a_rows=64, b_rows=64)), None)
    assert rc == ERR_ARG and "MN-major A with M=80" in msg


def test_every_status_is_negative_and_sticky_message_is_thread_local(abi):
    import threading

    capi, p, call = abi
    rc, msg = call("xva_attn_bin_loss", p, p, 0, 14, C.c_float(1e-12), p, None)
    assert rc < 0 and msg
    seen = {}

    def other():
        seen["msg"] = (capi.load().xva_last_error() or b"").decode()

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert seen["msg"] == ""                                                       # xva_last_error() is per thread
