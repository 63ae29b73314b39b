#!/bin/bash
# round 2, call P: xVAPitch decoder + VITS discriminator parity; the whole hifigan GPU file as regression
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hifigan_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2p_tests.log
cat gpurun_out/r2p_tests.log | cut -c1-1200
