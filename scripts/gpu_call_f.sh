# session-4 GPU call F: 16-row segments -- tests + A/B against 32-row minimum on the same box
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/f_tests.log
tail -3 gpurun_out/f_tests.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-hifigan"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['roofline']['achieved'],1), 'TF/s gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms gemm')" "$1" "$2" 2>&1 | tail -1; }
XVA_GEMM_MINSEG=32 timeout 300 python bench.py $B > gpurun_out/f_bench_seg32.log 2>&1; pick gpurun_out/f_bench_seg32.log seg32
XVA_BENCH_GEMM_TABLE=gpurun_out/f_fp_gemm_table.txt timeout 300 python bench.py $B > gpurun_out/f_bench_seg16.log 2>&1; pick gpurun_out/f_bench_seg16.log seg16
XVA_GEMM_MINSEG=32 timeout 300 python bench.py $B > gpurun_out/f_bench_seg32b.log 2>&1; pick gpurun_out/f_bench_seg32b.log seg32b
timeout 300 python bench.py $B > gpurun_out/f_bench_seg16b.log 2>&1; pick gpurun_out/f_bench_seg16b.log seg16b
XVA_GEMM_MINSEG=32 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/f_hifigan_seg32.log 2>&1
tail -1 gpurun_out/f_hifigan_seg32.log | cut -c1-160
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/f_hifigan.log 2>&1
tail -1 gpurun_out/f_hifigan.log | cut -c1-160
