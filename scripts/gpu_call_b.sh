# session-4 GPU call B: new kernel features (groups, segmented tiles) + restructured discriminators
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/b_gemm_tests.log
tail -3 gpurun_out/b_gemm_tests.log
(timeout 600 python -m pytest tests/test_hifigan_gpu.py -m gpu -q 2>&1 | tail -60) > gpurun_out/b_hifigan_tests.log
tail -3 gpurun_out/b_hifigan_tests.log
timeout 300 python scripts/diag_gen_grad.py > gpurun_out/b_diag_gen.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/b_tests.log
tail -3 gpurun_out/b_tests.log
XVA_BENCH_GEMM_TABLE=gpurun_out/b_fp_gemm_table.txt timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/b_bench.log 2>&1
tail -1 gpurun_out/b_bench.log | cut -c1-400
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/b_hifigan.log 2>&1
tail -1 gpurun_out/b_hifigan.log | cut -c1-400
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/b_hifigan_gemm_table.txt > gpurun_out/b_hifigan_prof.log 2>&1
head -3 gpurun_out/b_hifigan_gemm_table.txt
XVA_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/b_hifigan_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/b_hifigan_launches.csv gpurun_out/b_hifigan_launches_summary.txt "HiFi-GAN B=16x8192 training step, eager" | head -30
