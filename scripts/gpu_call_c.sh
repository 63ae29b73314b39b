# session-4 GPU call C: row kernels / LN un-fuse / wn_pack validation + same-box A/B against the session-start build
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/c_tests.log
tail -4 gpurun_out/c_tests.log
B="--steps 30 --warmup 5 --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['roofline']['achieved'],1), 'TF/s gemm', round(d['roofline']['kernel_ms_per_step'],2), 'ms gemm')" "$1" "$2" 2>&1 | tail -1; }
(cd old_tree && timeout 300 python bench.py $B) > gpurun_out/c_bench_old1.log 2>&1; pick gpurun_out/c_bench_old1.log old1
XVA_BENCH_GEMM_TABLE=gpurun_out/c_fp_gemm_table.txt timeout 300 python bench.py $B > gpurun_out/c_bench_new1.log 2>&1; pick gpurun_out/c_bench_new1.log new1
(cd old_tree && timeout 300 python bench.py $B) > gpurun_out/c_bench_old2.log 2>&1; pick gpurun_out/c_bench_old2.log old2
timeout 300 python bench.py $B > gpurun_out/c_bench_new2.log 2>&1; pick gpurun_out/c_bench_new2.log new2
XVA_FUSE_LN=1 timeout 300 python bench.py $B > gpurun_out/c_bench_fuseln.log 2>&1; pick gpurun_out/c_bench_fuseln.log fuse_ln
XVA_GEMM_SEG=0 XVA_GEMM_FILL=0 timeout 300 python bench.py $B > gpurun_out/c_bench_noseg.log 2>&1; pick gpurun_out/c_bench_noseg.log noseg_nofill
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/c_hifigan.log 2>&1
tail -1 gpurun_out/c_hifigan.log | cut -c1-300
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/c_hifigan_gemm_table.txt > gpurun_out/c_hifigan_prof.log 2>&1
head -2 gpurun_out/c_hifigan_gemm_table.txt
XVA_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/c_hifigan_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/c_hifigan_launches.csv gpurun_out/c_hifigan_launches_summary.txt "HiFi-GAN B=16x8192 training step, eager" | head -24
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_fp_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/c_fp_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/c_fp_launches.csv gpurun_out/c_fp_launches_summary.txt "FastPitch B=32x880 stage-3 step, eager, 3 warm-up + 1 timed + 1 e2e + 1 instrumented steps" | head -24
