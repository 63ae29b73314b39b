#!/bin/bash
# round 2, call V: one side stream per launching stream -- regression tests, then the HiFi-GAN / xVAPitch timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py tests/test_vits_gpu.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2v_tests.log; cut -c1-800 gpurun_out/r2v_tests.log
if grep -q "failed" gpurun_out/r2v_tests.log; then exit 1; fi
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.log 2>&1
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2v_bench.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]; x = d.get("xvapitch_hifi_only") or {}
print("fastpitch", round(d["ms_per_step"], 3), "hifigan", round(h["ms_per_step"], 3), "ms/step e2e", round(h["e2e"]["ms_per_step"], 3), "loss", h["loss_gen_all"],
      "| xvapitch", x.get("ms_per_step"), x.get("error"))
PY
