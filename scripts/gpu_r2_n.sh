# round 2, call N: 16-warp backward kernels (default) and the pipelined forward (XVA_ATTN_PIPE=4)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for pipe in 0 4; do
  XVA_ATTN_PIPE=$pipe timeout 120 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -s -x > gpurun_out/r2n_attn_tests_pipe$pipe.log 2>&1
  rc=$?
  echo "pipe=$pipe rc=$rc"; grep -E "attn_(fwd|bwd) B|passed|failed|Error" gpurun_out/r2n_attn_tests_pipe$pipe.log | cut -c1-200 | head -12
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2n_attn_tests_pipe$pipe.log | cut -c1-250; fi
done
if grep -q "passed" gpurun_out/r2n_attn_tests_pipe0.log && ! grep -q "failed" gpurun_out/r2n_attn_tests_pipe0.log; then
  pipes="0"
  if grep -q "passed" gpurun_out/r2n_attn_tests_pipe4.log && ! grep -q "failed" gpurun_out/r2n_attn_tests_pipe4.log; then pipes="0 4"; fi
  for pipe in $pipes; do
    XVA_ATTN_PIPE=$pipe timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2n_bench_pipe$pipe.log 2>&1
    python - $pipe <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2n_bench_pipe{tag}.log").read().splitlines() if l.startswith("{")][-1])
    a = d["roofline"].get("attention", {})
    print("pipe", tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s | attention", round(a.get("kernel_ms_per_step", 0), 3), "ms/step frac", round(a.get("frac", 0), 4), "| gemm frac", round(d["roofline"]["frac"], 4), "loss", d["loss"])
except Exception as e:
    print(tag, "failed", e); print(open(f"gpurun_out/r2n_bench_pipe{tag}.log").read()[-1500:])
PY
  done
fi
