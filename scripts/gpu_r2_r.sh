#!/bin/bash
# round 2, call R: hifi_only step (device-side segments) tests; xvapitch bench entry alone (graph + eager); full default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vits_gpu.py -q -m gpu 2>&1 | tail -12 > gpurun_out/r2r_vits.log
cut -c1-1200 gpurun_out/r2r_vits.log
timeout 400 python bench.py --xvapitch-only > gpurun_out/r2r_xva_graph.log 2>&1; tail -c 2500 gpurun_out/r2r_xva_graph.log
timeout 400 python bench.py --xvapitch-only --no-graph > gpurun_out/r2r_xva_eager.log 2>&1; tail -c 600 gpurun_out/r2r_xva_eager.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2r_bench.log 2>&1; tail -c 6000 gpurun_out/r2r_bench.log
