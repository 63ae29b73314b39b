# round 2, call J: evidence -- ncu launch list of the bench step, ncu --set full of the dominant tap-GEMM launch, compute-sanitizer
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python scripts/diag_batch_invariance.py > gpurun_out/r2j_batch_invariance.txt 2>&1; grep -v "0.000e+00" gpurun_out/r2j_batch_invariance.txt | head -30
XVA_FUSED_ATTN=0 timeout 120 python scripts/diag_batch_invariance.py > gpurun_out/r2j_batch_invariance_unfused.txt 2>&1; grep -v "0.000e+00" gpurun_out/r2j_batch_invariance_unfused.txt | head -30
# (1) launch list of the bench command (eager launches: one ncu record per kernel), FastPitch half then HiFi-GAN half
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2j_fp_launches.csv python bench.py --no-hifigan --no-cpu-baseline --no-graph --steps 2 --warmup 3 > gpurun_out/r2j_fp_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2j_fp_launches.csv gpurun_out/r2j_fp_launches_summary.txt "FastPitch B=32x880 stage-3 step, eager, bench.py --no-graph --steps 2 --warmup 3 (3 warm-up + 2 timed + 2 e2e + 2 eager instrumented-pass steps)" | head -25
# (2) the dominant launch (ConvFF second conv, N = 384, K = 3 x 1536) under --set full
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2j_conv2 python scripts/prof_gemm.py conv2 3 > gpurun_out/r2j_conv2_ncu.log 2>&1
ls -la gpurun_out/r2j_conv2.ncu-rep
# (3) compute-sanitizer: memcheck over the GEMM / attention / MAS / stage-1 kernel tests, racecheck over the shared-memory heavy ones
SAN="compute-sanitizer --error-exitcode 7 --print-limit 20"
(timeout 420 $SAN --tool memcheck python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "not full_size" 2>&1 | tail -25) > gpurun_out/r2j_memcheck_gemm.log; tail -4 gpurun_out/r2j_memcheck_gemm.log
(timeout 300 $SAN --tool memcheck python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -x -k "not speed" 2>&1 | tail -25) > gpurun_out/r2j_memcheck_attn.log; tail -4 gpurun_out/r2j_memcheck_attn.log
(timeout 300 $SAN --tool memcheck python -m pytest tests/test_mas_gpu.py tests/test_stage1_gpu.py tests/test_regulate_gpu.py -m gpu -q -x -k "not full_size and not properties" 2>&1 | tail -25) > gpurun_out/r2j_memcheck_stage1.log; tail -4 gpurun_out/r2j_memcheck_stage1.log
(timeout 300 $SAN --tool racecheck python -m pytest tests/test_mas_gpu.py tests/test_attn_fused_gpu.py -m gpu -q -x -k "(golden or maximum_path or (matches_fp64 and 160)) and not speed" 2>&1 | tail -25) > gpurun_out/r2j_racecheck.log; tail -4 gpurun_out/r2j_racecheck.log
