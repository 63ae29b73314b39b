mkdir -p gpurun_out
for k in conv1 conv2ln dgrad2 wgrad1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_$k python scripts/prof_gemm.py $k 3 > gpurun_out/prof_$k.log 2>&1
  tail -2 gpurun_out/prof_$k.log
done
ls -la gpurun_out/*.ncu-rep
