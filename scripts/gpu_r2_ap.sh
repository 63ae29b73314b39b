#!/bin/bash
# round 2, call AP: WaveNet backward with its weight gradients on the side stream -- tests (eager, streams forced, flows), then A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vits_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2ap_tests.log; cut -c1-500 gpurun_out/r2ap_tests.log
if grep -q "failed\|error" gpurun_out/r2ap_tests.log; then exit 1; fi
for v in auto 0 auto 0; do
  XVA_BWD_STREAMS=$v timeout 300 python bench.py --xvapitch-only --hifigan-steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith(chr(123))][-1]); print('XVA_BWD_STREAMS=$v xvapitch', round(d['ms_per_step'],3))"
done
