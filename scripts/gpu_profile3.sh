# profiles for the round: launch list of one full step + ncu --set full of the dominant GEMM shapes
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1050 -c 360 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r01.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki][:70]][0] += 1; agg[r[ki][:70]][1] += float(r[vi].replace(',', ''))
    except Exception: pass
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/launches_r01_summary.txt', 'w') as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)-1} launches, total {tot/1e6:.3f} ms (cold-cache, serialised)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:70s} n={v[0]:4d} {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%\n")
print(open('gpurun_out/launches_r01_summary.txt').read())
PY
for k in conv1 conv2ln dgrad2 wgrad1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r01_$k python scripts/prof_gemm.py $k 3 > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out/r01_*.ncu-rep
