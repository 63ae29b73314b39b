# round 2, call F: fused attention (forward + backward) -- tests first, everything else only if they pass
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -x -k "not speed" > gpurun_out/r2f_attn_tests.log 2>&1
rc=$?
tail -25 gpurun_out/r2f_attn_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "ATTENTION TESTS rc=$rc: stopping"; exit 0; fi
true
timeout 600 python -m pytest tests/test_fastpitch_gpu.py tests/test_parity_full_gpu.py tests/test_infer_gpu.py -m gpu -q -x > gpurun_out/r2f_fp_tests.log 2>&1
rc=$?
tail -15 gpurun_out/r2f_fp_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "FASTPITCH TESTS rc=$rc: stopping"; exit 0; fi
for f in 0 1; do
  XVA_FUSED_ATTN=$f XVA_BENCH_GEMM_TABLE=gpurun_out/r2f_table_attn$f.txt timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2f_bench_attn$f.log 2>&1
done
python - <<'PY'
import json
for tag in ("attn0", "attn1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2f_bench_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms_per_step"], 3), "ms frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"] // d["steps"], "loss", d["loss"])
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2f_bench_{tag}.log").read()[-1500:])
PY
timeout 200 python -m pytest tests/test_trainers_gpu.py -m gpu -q -x > gpurun_out/r2f_trainer_tests.log 2>&1; tail -8 gpurun_out/r2f_trainer_tests.log | cut -c1-300
cat > /tmp/attn_prof.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import ops
B, T = 32, 880
qkv = torch.randn(B, T, 192, device="cuda")
ops.round_tf32_(qkv.view(-1), qkv.view(-1))
lens = torch.full((B,), T, device="cuda", dtype=torch.int32)
sd = torch.zeros(1, device="cuda", dtype=torch.int64)
dvec = torch.randn(B, T, 64, device="cuda")
for _ in range(3):
    vec, lse = ops.attn_fwd(qkv, lens, 0.125, 0.1, 5, sd)
    ops.attn_bwd(qkv, dvec, vec, lse, lens, 0.125, 0.1, 5, sd)
torch.cuda.synchronize()
PY
for k in attn_fwd attn_bwd_dq attn_bwd_dkv; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o gpurun_out/r2f_${k} python /tmp/attn_prof.py > gpurun_out/r2f_ncu_${k}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
