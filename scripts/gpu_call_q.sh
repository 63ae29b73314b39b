mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/q_gemm_tests.log
tail -4 gpurun_out/q_gemm_tests.log
if grep -q "passed" gpurun_out/q_gemm_tests.log && ! grep -q "failed" gpurun_out/q_gemm_tests.log; then
  (timeout 600 python -m pytest tests/test_hifigan_gpu.py tests/test_fastpitch_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/q_tests.log
  tail -3 gpurun_out/q_tests.log
  XVA_GEMM_MTAP=0 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/q_hifigan_nomtap.log 2>&1; tail -1 gpurun_out/q_hifigan_nomtap.log | cut -c1-150
  timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/q_hifigan.log 2>&1; tail -1 gpurun_out/q_hifigan.log | cut -c1-150
  timeout 600 python scripts/bench_generator_large.py 8 880 gpurun_out/q_generator_large_table.txt > gpurun_out/q_gen_large.log 2>&1
  head -14 gpurun_out/q_generator_large_table.txt | cut -c1-200
fi
