# session-4 GPU call D: ncu --set full of the epilogue-bound launches + the new bench line
mkdir -p gpurun_out
for k in qk onet dgrad2 conv2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/d_$k python scripts/prof_gemm.py $k 3 > gpurun_out/d_prof_$k.log 2>&1
done
ls -la gpurun_out/d_*.ncu-rep
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/d_bench.log 2>&1
tail -1 gpurun_out/d_bench.log | cut -c1-1500
