#!/bin/bash
# round 2, the last GPU call: smoke() with the text-encoder check, then the bench's next-tier child (hifi_only step + text encoder)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 18 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2as_smoke.log 2>&1; tail -4 gpurun_out/r2as_smoke.log | cut -c1-600
( time timeout 24 python bench.py --xvapitch-only --hifigan-steps 5 ) > gpurun_out/r2as_bench_child.log 2>&1; tail -c 2500 gpurun_out/r2as_bench_child.log
