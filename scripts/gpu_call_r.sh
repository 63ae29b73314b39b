# session-4 GPU call R: final validation of the round -- full GPU suite, smoke(), the bench line, launch lists, one ncu --set full
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r_tests.log
tail -3 gpurun_out/r_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r_smoke.log 2>&1; tail -1 gpurun_out/r_smoke.log | cut -c1-300
XVA_BENCH_GEMM_TABLE=gpurun_out/r_fp_gemm_table.txt timeout 600 python bench.py > gpurun_out/r_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_bench.log').read().strip().splitlines()[-1])
print('fastpitch', round(d['ms_per_step'],3), 'ms', round(d['value']), 'frames/s; gemm', round(d['roofline']['achieved'],1), 'TF/s frac', round(d['roofline']['frac'],3), '; e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
print('dominant', d['roofline']['dominant_launch'])
h=d['hifigan']; print('hifigan', round(h['ms_per_step'],2), 'ms', round(h['value']), 'samples/s; gemm', round(h['roofline']['achieved'],1), 'TF/s; launches', h['gpu_launches_per_step'], 'cpu', h.get('cpu_baseline',{}).get('value'))
print('cpu', d['cpu_baseline'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_fp_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-hifigan > gpurun_out/r_fp_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r_fp_launches.csv gpurun_out/r_fp_launches_summary.txt "FastPitch B=32x880 stage-3 step, eager, 3 warm-up + 1 timed + 1 e2e + 1 instrumented steps (6 steps)" | head -12
XVA_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/r_hifigan_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r_hifigan_launches.csv gpurun_out/r_hifigan_launches_summary.txt "HiFi-GAN B=16x8192 training step, eager" | head -10
for k in conv1 conv2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r_$k python scripts/prof_gemm.py $k 3 > gpurun_out/r_prof_$k.log 2>&1
done
ls -la gpurun_out/r_*.ncu-rep
