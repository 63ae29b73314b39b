# round 2, call M: software-pipelined attention backward kernels (XVA_ATTN_PIPE: bit 0 = dQ, bit 1 = dK/dV)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for pipe in 1 2; do
  XVA_ATTN_PIPE=$pipe timeout 120 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -s -x -k "bwd" > gpurun_out/r2m_attn_tests_pipe$pipe.log 2>&1
  rc=$?
  echo "pipe=$pipe rc=$rc"; grep -E "attn_bwd|passed|failed|Error" gpurun_out/r2m_attn_tests_pipe$pipe.log | cut -c1-200 | head -12
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2m_attn_tests_pipe$pipe.log | cut -c1-250; fi
done
if grep -q "passed" gpurun_out/r2m_attn_tests_pipe1.log && grep -q "passed" gpurun_out/r2m_attn_tests_pipe2.log && ! grep -q "failed" gpurun_out/r2m_attn_tests_pipe1.log gpurun_out/r2m_attn_tests_pipe2.log; then
  for pipe in 0 3; do
    XVA_ATTN_PIPE=$pipe timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2m_bench_pipe$pipe.log 2>&1
    python - $pipe <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2m_bench_pipe{tag}.log").read().splitlines() if l.startswith("{")][-1])
    a = d["roofline"].get("attention", {})
    print("pipe", tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s | attention", round(a.get("kernel_ms_per_step", 0), 3), "ms/step frac", round(a.get("frac", 0), 4), "| gemm frac", round(d["roofline"]["frac"], 4), "loss", d["loss"])
except Exception as e:
    print(tag, "failed", e); print(open(f"gpurun_out/r2m_bench_pipe{tag}.log").read()[-1500:])
PY
  done
fi
