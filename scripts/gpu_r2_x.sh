#!/bin/bash
# round 2, call X (2 GPUs): the whole GPU suite incl. the 2-rank tests, smoke, bench at N=2 and N=1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_gpu_tests.log 2>&1; tail -6 gpurun_out/r2x_gpu_tests.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; tail -1 gpurun_out/r2x_smoke.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n2.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n1.log 2>&1
python - <<'PY'
import json
for tag in ("bench_n2", "bench_n1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2x_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        h = d.get("hifigan") or {}; x = d.get("xvapitch_hifi_only") or {}
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| hifigan", round(h.get("ms_per_step", 0), 2), round(h.get("value", 0)),
              "| xva", x.get("ms_per_step"), x.get("error"), "| frac", (d.get("roofline") or {}).get("frac"), "clocks", d.get("clocks"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2x_{tag}.log").read()[-1500:])
PY
