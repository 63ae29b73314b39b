#!/bin/bash
# round 2, call S: where the xVAPitch hifi_only step spends its time (per-shape GEMM table + ncu launch list)
mkdir -p gpurun_out
timeout 300 python scripts/prof_vits.py gpurun_out/r2s_vits_gemm_table.txt > gpurun_out/r2s_table.log 2>&1; head -45 gpurun_out/r2s_vits_gemm_table.txt
XVA_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_vits_launches.csv python scripts/prof_vits.py > gpurun_out/r2s_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2s_vits_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0; n = 0
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    e = agg.setdefault(r[ki], [0, 0.0]); e[0] += 1; e[1] += v; tot += v; n += 1
out = [f"xVAPitch --hifi_only step, batch 16 x 256 frames, eager; ncu --metrics gpu__time_duration.sum --clock-control none, {n} launches, total {tot/1e3:.3f} ms (cold-cache, serialised)"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    out.append(f"{k[:90]:90s} n={c:5d} {t:10.1f} us {100*t/tot:5.1f}%")
open("gpurun_out/r2s_vits_launches_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[:32]))
PY
