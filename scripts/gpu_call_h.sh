# session-4 GPU call H: does a 384/512-column single-buffer tile beat 192/256-column double-buffered tiles on the L2-bound shapes?
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-hifigan"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['roofline']['achieved'],1), 'TF/s gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms gemm')" "$1" "$2" 2>&1 | tail -1; }
XVA_BENCH_GEMM_TABLE=gpurun_out/h_table_default.txt timeout 300 python bench.py $B > gpurun_out/h_bench_default.log 2>&1; pick gpurun_out/h_bench_default.log default
XVA_GEMM_NTILE=512 XVA_BENCH_GEMM_TABLE=gpurun_out/h_table_nt512.txt timeout 300 python bench.py $B > gpurun_out/h_bench_nt512.log 2>&1; pick gpurun_out/h_bench_nt512.log nt512
XVA_GEMM_NTILE=384 XVA_BENCH_GEMM_TABLE=gpurun_out/h_table_nt384.txt timeout 300 python bench.py $B > gpurun_out/h_bench_nt384.log 2>&1; pick gpurun_out/h_bench_nt384.log nt384
XVA_GEMM_PAIR=0 XVA_BENCH_GEMM_TABLE=gpurun_out/h_table_nopair.txt timeout 300 python bench.py $B > gpurun_out/h_bench_nopair.log 2>&1; pick gpurun_out/h_bench_nopair.log nopair
