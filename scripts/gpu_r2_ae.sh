#!/bin/bash
# round 2, call AE: predictors next to the decoder (FastPitch) -- tests, then A/B inside one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fastpitch_gpu.py tests/test_parity_full_gpu.py tests/test_trainers_gpu.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2ae_tests.log; cut -c1-700 gpurun_out/r2ae_tests.log
if grep -q "failed\|error" gpurun_out/r2ae_tests.log; then exit 1; fi
for v in 1 0 1 0; do
  XVA_PRED_STREAM=$v timeout 400 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-hifigan > gpurun_out/r2ae_bench_$v.log 2>&1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2ae_bench_$v.log").read().splitlines() if l.startswith("{")][-1])
print("pred_stream=$v fastpitch", round(d["ms_per_step"], 3), "ms/step e2e", round(d["e2e"]["ms_per_step"], 3), "loss", d["loss"])
PY
done
