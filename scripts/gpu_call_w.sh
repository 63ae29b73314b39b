# session-5 GPU call W: final validation of the round -- full GPU suite, smoke(), the bench line, launch lists of the three
# steps (FastPitch stage 3, HiFi-GAN, FastPitch stage 1), ncu --set full of the dominant GEMM and the two sequential kernels
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/w_tests.log
tail -3 gpurun_out/w_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; tail -1 gpurun_out/w_smoke.log | cut -c1-420
XVA_BENCH_GEMM_TABLE=gpurun_out/w_fp_gemm_table.txt timeout 600 python bench.py > gpurun_out/w_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench.log').read().strip().splitlines()[-1])
print('fastpitch', round(d['ms_per_step'],3), 'ms', round(d['value']), 'frames/s; gemm', round(d['roofline']['achieved'],1), 'TF/s frac', round(d['roofline']['frac'],3), '; e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
print('dominant', d['roofline']['dominant_launch'])
h=d['hifigan']; print('hifigan', round(h['ms_per_step'],2), 'ms', round(h['value']), 'samples/s; gemm', round(h['roofline']['achieved'],1), 'TF/s; launches', h['gpu_launches_per_step'], 'cpu', h.get('cpu_baseline',{}).get('value'))
print('cpu', d['cpu_baseline'])
PY
timeout 300 python scripts/bench_stage1.py 20 > gpurun_out/w_stage1_bench.log 2>&1; tail -1 gpurun_out/w_stage1_bench.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_fp_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-hifigan > gpurun_out/w_fp_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/w_fp_launches.csv gpurun_out/w_fp_launches_summary.txt "FastPitch B=32x880 stage-3 step, eager, 3 warm-up + 1 timed + 1 e2e + 1 instrumented steps (6 steps)" | head -8
XVA_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/w_hifigan_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/w_hifigan_launches.csv gpurun_out/w_hifigan_launches_summary.txt "HiFi-GAN B=16x8192 training step, eager" | head -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_stage1_launches.csv python scripts/bench_stage1.py 1 --no-cpu > gpurun_out/w_stage1_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/w_stage1_launches.csv gpurun_out/w_stage1_launches_summary.txt "FastPitch stage-1 step B=32x880x160, eager, 3 warm-up + 1 timed + 1 instrumented steps (5 steps)" | head -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/w_conv2 python scripts/prof_gemm.py conv2 3 > gpurun_out/w_prof_conv2.log 2>&1
for k in ctc_recursion mas; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o gpurun_out/w_$k python scripts/bench_stage1.py 1 --no-cpu > gpurun_out/w_prof_$k.log 2>&1
done
ls -la gpurun_out/w_*.ncu-rep
