# session-4 GPU call E: restructured epilogue -- tests, A/B against the previous build on the same box
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/e_tests.log
tail -3 gpurun_out/e_tests.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-hifigan"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['roofline']['achieved'],1), 'TF/s gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms gemm')" "$1" "$2" 2>&1 | tail -1; }
(cd old_tree && timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline) > gpurun_out/e_bench_old1.log 2>&1; pick gpurun_out/e_bench_old1.log old1
XVA_BENCH_GEMM_TABLE=gpurun_out/e_fp_gemm_table.txt timeout 300 python bench.py $B > gpurun_out/e_bench_new1.log 2>&1; pick gpurun_out/e_bench_new1.log new1
timeout 300 python bench.py $B > gpurun_out/e_bench_new2.log 2>&1; pick gpurun_out/e_bench_new2.log new2
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/e_hifigan.log 2>&1
tail -1 gpurun_out/e_hifigan.log | cut -c1-200
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/e_hifigan_gemm_table.txt > gpurun_out/e_hifigan_prof.log 2>&1
head -2 gpurun_out/e_hifigan_gemm_table.txt
for k in qk onet; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/e_$k python scripts/prof_gemm.py $k 3 > gpurun_out/e_prof_$k.log 2>&1
done
