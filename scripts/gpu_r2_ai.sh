#!/bin/bash
# round 2, call AI: ncu --set full of the HiFi-GAN small-channel regime (32 channels, kernel 11, 16 x 8192): forward and weight gradient
mkdir -p gpurun_out
for w in hg32 hg32w; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2ai_$w python scripts/prof_gemm.py $w 3 > gpurun_out/r2ai_${w}_ncu.log 2>&1
  ncu -i gpurun_out/r2ai_$w.ncu-rep --page raw --csv > gpurun_out/r2ai_$w.csv 2>/dev/null
  python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2ai_$w.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for k in want:
    if k in hdr:
        i = hdr.index(k); out.append(f"{k:75s} {vals[i]} {units[i]}")
open("gpurun_out/r2ai_${w}_summary.txt", "w").write("\n".join(out) + "\n")
print("== $w"); print("\n".join(out))
PY
done
