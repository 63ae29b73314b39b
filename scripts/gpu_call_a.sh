# session-4 GPU call A: full GPU test suite, FastPitch bench with per-shape table, HiFi-GAN bench + where-the-time-goes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/a_tests.log 2>&1
tail -5 gpurun_out/a_tests.log
XVA_BENCH_GEMM_TABLE=gpurun_out/a_fp_gemm_table.txt timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/a_bench.log 2>&1
tail -1 gpurun_out/a_bench.log | cut -c1-3000
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/a_hifigan.log 2>&1
tail -1 gpurun_out/a_hifigan.log
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/a_hifigan_gemm_table.txt > gpurun_out/a_hifigan_prof.log 2>&1
head -3 gpurun_out/a_hifigan_gemm_table.txt
XVA_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_hifigan_launches.csv python scripts/prof_hifigan.py > gpurun_out/a_hifigan_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/a_hifigan_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki][:90]][0] += 1; agg[r[ki][:90]][1] += float(r[vi].replace(',', ''))
    except Exception: pass
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/a_hifigan_launches_summary.txt', 'w') as f:
    f.write(f"HiFi-GAN B=16x8192 training step, eager; ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)-1} launches, total {tot/1e6:.3f} ms (cold-cache, serialised)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:90s} n={v[0]:5d} {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%\n")
print(open('gpurun_out/a_hifigan_launches_summary.txt').read()[:5000])
PY
rm -f gpurun_out/a_hifigan_launches.csv.tmp
