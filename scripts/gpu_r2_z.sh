#!/bin/bash
# round 2, call Z: aligned AdamW arena + vectorised pack -- tests, then A/B of the pack path inside one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py tests/test_vits_gpu.py tests/test_trainers_gpu.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2z_tests.log; cut -c1-600 gpurun_out/r2z_tests.log
if grep -q "failed\|error" gpurun_out/r2z_tests.log; then exit 1; fi
timeout 200 python scripts/bench_wnpack.py 2>&1 | tail -2
for v in 1 0 1 0; do
  XVA_WNPACK_VEC=$v timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --hifigan-steps 30 --no-cpu-baseline > gpurun_out/r2z_bench_$v.log 2>&1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2z_bench_$v.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]; x = d.get("xvapitch_hifi_only") or {}
print("vec=$v hifigan", round(h["ms_per_step"], 3), "| xvapitch", x.get("ms_per_step"), x.get("error"))
PY
done
