#!/bin/bash
# round 2, call Q: xVAPitch hifi_only path parity (WN, posterior encoder, mel, full step) + hifigan regression
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vits_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2q_vits.log
cut -c1-1500 gpurun_out/r2q_vits.log
timeout 600 python -m pytest tests/test_hifigan_gpu.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2q_hifigan.log
cut -c1-600 gpurun_out/r2q_hifigan.log
