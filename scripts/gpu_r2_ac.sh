#!/bin/bash
# round 2, call AC: tiled first-layer kernels -- discriminator tests (plain + memcheck), A/B inside one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py tests/test_vits_gpu.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2ac_tests.log; cut -c1-900 gpurun_out/r2ac_tests.log
if grep -q "failed\|error" gpurun_out/r2ac_tests.log; then exit 1; fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_hifigan_gpu.py -q -m gpu -k "discriminator_forward or discriminator_step or vits_discriminator" > gpurun_out/r2ac_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/r2ac_sanitizer.log | cut -c1-200
for v in 1 0 1 0; do
  XVA_C1_TILED=$v timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --hifigan-steps 30 --no-cpu-baseline > gpurun_out/r2ac_bench_$v.log 2>&1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2ac_bench_$v.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]; x = d.get("xvapitch_hifi_only") or {}
print("c1_tiled=$v hifigan", round(h["ms_per_step"], 3), "| xvapitch", x.get("ms_per_step"), x.get("error"))
PY
done
