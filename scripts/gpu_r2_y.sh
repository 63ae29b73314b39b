#!/bin/bash
# round 2, call Y: vectorised weight pack / unpack -- packer tests (plain + under compute-sanitizer), model tests, timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hifigan_gpu.py tests/test_vits_gpu.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2y_tests.log; cut -c1-600 gpurun_out/r2y_tests.log
if grep -q "failed\|error" gpurun_out/r2y_tests.log; then exit 1; fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_hifigan_gpu.py -q -m gpu -k "packer or xvapitch_decoder_keys or state_dict" > gpurun_out/r2y_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2y_sanitizer.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.log 2>&1
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2y_bench.log").read().splitlines() if l.startswith("{")][-1])
h = d["hifigan"]; x = d.get("xvapitch_hifi_only") or {}
print("fastpitch", round(d["ms_per_step"], 3), "hifigan", round(h["ms_per_step"], 3), "| xvapitch", x.get("ms_per_step"), x.get("error"))
PY
