# one-GPU profiling pass: per-shape GEMM table, ncu launch list of 1 step, ncu --set full of the ConvFF GEMMs
mkdir -p gpurun_out
XVA_BENCH_GEMM_TABLE=gpurun_out/gemm_table.txt python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tab.log 2>&1
cat gpurun_out/gemm_table.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki][:60]][0] += 1; agg[r[ki][:60]][1] += float(r[vi].replace(',', ''))
    except Exception: pass
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={v[0]:4d} {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%")
PY
