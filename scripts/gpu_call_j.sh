# session-4 GPU call J: MAS kernel, fused softmax backward, row-tile pair rule
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/j_tests.log
tail -5 gpurun_out/j_tests.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-hifigan"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['roofline']['achieved'],1), 'TF/s gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms gemm')" "$1" "$2" 2>&1 | tail -1; }
XVA_FUSE_SOFTMAX_BWD=0 timeout 300 python bench.py $B > gpurun_out/j_bench_unfused.log 2>&1; pick gpurun_out/j_bench_unfused.log softmax_bwd_unfused
XVA_BENCH_GEMM_TABLE=gpurun_out/j_fp_gemm_table.txt timeout 300 python bench.py $B > gpurun_out/j_bench_fused.log 2>&1; pick gpurun_out/j_bench_fused.log softmax_bwd_fused
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/j_hifigan.log 2>&1
tail -1 gpurun_out/j_hifigan.log | cut -c1-160
