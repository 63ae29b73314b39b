#!/bin/bash
# round 2, call AA: native spectral-norm pack -- its own test, discriminator / step tests, sanitizer, A/B inside one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py tests/test_trainers_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2aa_tests.log; cut -c1-1500 gpurun_out/r2aa_tests.log
if grep -q "failed\|error" gpurun_out/r2aa_tests.log; then exit 1; fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_hifigan_gpu.py -q -m gpu -k "spectral_packer or discriminator_forward" > gpurun_out/r2aa_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/r2aa_sanitizer.log | cut -c1-200
for v in 1 0 1 0; do
  XVA_SN_NATIVE=$v timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --hifigan-steps 30 --no-cpu-baseline --no-xvapitch > gpurun_out/r2aa_bench_$v.log 2>&1
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2aa_bench_$v.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]
print("sn_native=$v hifigan", round(h["ms_per_step"], 3), "launches", h["gpu_launches_per_step"], "loss", h["loss_gen_all"])
PY
done
