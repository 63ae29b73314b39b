"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
src, dst, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki][:90]][0] += 1; agg[r[ki][:90]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
tot = sum(v[1] for v in agg.values())
with open(dst, 'w') as f:
    f.write(f"{title}; ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows) - 1} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:90s} n={v[0]:5d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%\n")
print(open(dst).read())
