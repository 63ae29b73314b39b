#!/bin/bash
# round 2, last GPU seconds: the text-encoder GPU tests alone (never run on hardware before this call)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_vits_text_encoder_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2aq_textenc.log 2>&1
tail -25 gpurun_out/r2aq_textenc.log | cut -c1-400
