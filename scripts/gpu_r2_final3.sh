#!/bin/bash
# round 2, final check on one GPU: the whole GPU suite, smoke(), the driver's bench command, the reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f3_gpu_tests.log 2>&1; tail -4 gpurun_out/r2f3_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f3_smoke.log 2>&1; tail -1 gpurun_out/r2f3_smoke.log | cut -c1-500
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2f3_bench_n1.log 2> gpurun_out/r2f3_bench_n1.err; tail -3 gpurun_out/r2f3_bench_n1.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2f3_bench_ref.log 2> gpurun_out/r2f3_bench_ref.err; tail -3 gpurun_out/r2f3_bench_ref.err
python - <<'PY'
import json
for tag in ("bench_n1", "bench_ref"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2f3_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        h = d.get("hifigan") or {}; x = d.get("xvapitch_hifi_only") or {}; r = d.get("roofline") or {}
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "| hifigan", round(h.get("ms_per_step", 0), 2), round(h.get("value", 0)),
              "| xva", round(x.get("ms_per_step", 0), 2), round(x.get("value", 0)), x.get("error"), "| frac", r.get("frac"), "hifi frac", (h.get("roofline") or {}).get("frac"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2f3_{tag}.log").read()[-1500:])
PY
