#!/bin/bash
# round 2, call U: MPD and MSD side by side -- regression tests, then the HiFi-GAN half of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py tests/test_trainers_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2u_tests.log; cut -c1-1500 gpurun_out/r2u_tests.log
if grep -q "failed" gpurun_out/r2u_tests.log; then exit 1; fi
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-xvapitch > gpurun_out/r2u_bench.log 2>&1
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2u_bench.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]
print("fastpitch", round(d["ms_per_step"], 3), "hifigan", round(h["ms_per_step"], 3), "ms/step e2e", round(h["e2e"]["ms_per_step"], 3), "launches", h["gpu_launches_per_step"], "loss", h["loss_gen_all"])
PY
