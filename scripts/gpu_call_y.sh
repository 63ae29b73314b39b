# session-5 GPU call Y (the round's last seconds): parity of the gated two-stream backward experiments
mkdir -p gpurun_out
(XVA_TEST_EXPERIMENTAL=1 timeout 28 python -m pytest tests/test_fastpitch_gpu.py tests/test_hifigan_gpu.py -m gpu -q -x -k two_stream 2>&1 | tail -15) > gpurun_out/y_streams.log
tail -6 gpurun_out/y_streams.log
