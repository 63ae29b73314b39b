# session-5 GPU call X (last GPU seconds of the round): FastPitch.infer parity
mkdir -p gpurun_out
(timeout 80 python -m pytest tests/test_infer_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/x_infer.log
tail -25 gpurun_out/x_infer.log
