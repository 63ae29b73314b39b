#!/bin/bash
# round 2, last GPU seconds: LayerNorm kernels extended to 1024 channels (the xVAPitch pitch predictor's 708 / 780)
mkdir -p gpurun_out
timeout 14 python -m pytest tests/test_rowops_gpu.py -m gpu -q -p no:cacheprovider -k "layernorm" > gpurun_out/r2at_ln.log 2>&1
tail -12 gpurun_out/r2at_ln.log | cut -c1-300
