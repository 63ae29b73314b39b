# session-4 GPU call I: per-shape pair rule, 384-column single tile, tiled conv_c1 kernels
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/i_tests.log
tail -3 gpurun_out/i_tests.log
XVA_BENCH_GEMM_TABLE=gpurun_out/i_fp_gemm_table.txt timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/i_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/i_bench.log').read().strip().splitlines()[-1])
print('fastpitch', round(d['ms_per_step'],3), 'ms', round(d['value']), 'frames/s; gemm', round(d['roofline']['achieved'],1), 'TF/s', round(d['roofline']['kernel_ms_per_step'],2), 'ms; e2e', round(d['e2e']['value']))
h=d['hifigan']; print('hifigan', round(h['ms_per_step'],2), 'ms', round(h['value']), 'samples/s; gemm', round(h['roofline']['achieved'],1), 'TF/s', round(h['roofline']['kernel_ms_per_step'],2), 'ms')
PY
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/i_hifigan_gemm_table.txt > gpurun_out/i_hifigan_prof.log 2>&1
head -2 gpurun_out/i_hifigan_gemm_table.txt
XVA_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/i_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/i_hifigan_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/i_hifigan_launches.csv gpurun_out/i_hifigan_launches_summary.txt "HiFi-GAN B=16x8192 training step, eager" | head -16
