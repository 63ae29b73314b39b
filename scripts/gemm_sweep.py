"""Times one tap-GEMM shape under the bring-up knobs of gemm_tc.cu (XVA_GEMM_DBG / XVA_GEMM_PAIR / XVA_GEMM_STAGES)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import __graft_entry__ as ge
    ge.build()
    from xva_trainer_b200 import ops
    B, T = 32, 880
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, h = r(B, T, 384), r(B, T, 1536)
    w1, w2 = r(3, 1536, 384) * 0.03, r(3, 384, 1536) * 0.02
    b1 = r(1536)
    K3 = (-1, 0, 1)
    out = torch.empty(B, T, 1536, device="cuda")
    def run(): ops.conv_fwd(x, w1, K3, bias=b1, relu=True, out=out)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    import ctypes
    from xva_trainer_b200 import capi
    buf = (ctypes.c_longlong * 8)()
    capi.call("xva_gemm_debug_counters", ctypes.byref(buf))
    names = ["mma wait acc", "mma wait ops", "mma issue", "epi wait", "epi work", "tiles", "total"]
    print(f"{os.environ.get('TAG','')}: {ms*1e3:7.1f} us  {2*B*T*1536*384*3/ms/1e9:6.1f} TFLOP/s  | " + ", ".join(f"{n}={buf[i]}" for i, n in enumerate(names)))
else:
    for tag, env in [("pair base", {"XVA_GEMM_DBG": "32"}), ("1cta base", {"XVA_GEMM_PAIR": "0", "XVA_GEMM_DBG": "32"}),
                     ("1cta raw-mma", {"XVA_GEMM_PAIR": "0", "XVA_GEMM_DBG": "49"}),
                     ("1cta no-store", {"XVA_GEMM_PAIR": "0", "XVA_GEMM_DBG": "33"}),
                     ]:
        e = dict(os.environ); e.update(env); e["TAG"] = tag
        subprocess.run([sys.executable, __file__, "child"], env=e)
