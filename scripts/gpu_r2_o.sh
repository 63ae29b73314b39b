#!/bin/bash
# round 2, call O: xVAPitch waveform decoder parity + bench clock-sampler check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hifigan_gpu.py -x -q -m gpu -k "xvapitch or packer or golden" 2>&1 | tail -15 > gpurun_out/r2o_tests.log
cat gpurun_out/r2o_tests.log
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench.log 2>&1
tail -c 3000 gpurun_out/r2o_bench.log
