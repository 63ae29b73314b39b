# session-5 GPU call T: stage-1 step timing at 32 x 880 x 160, launch list, ncu --set full of the score and CTC kernels
mkdir -p gpurun_out
timeout 300 python scripts/bench_stage1.py 20 > gpurun_out/t_stage1_bench.log 2>&1; tail -1 gpurun_out/t_stage1_bench.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_stage1_launches.csv python scripts/bench_stage1.py 1 --no-cpu > gpurun_out/t_stage1_ncu.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/t_stage1_launches.csv gpurun_out/t_stage1_launches_summary.txt "FastPitch stage-1 step B=32x880x160, eager, 3 warm-up + 1 timed + 1 instrumented steps" | head -14
for k in attn_score_fwd attn_ctc attn_score_bwd attn_key_grad; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o gpurun_out/t_$k python scripts/bench_stage1.py 1 --no-cpu > gpurun_out/t_prof_$k.log 2>&1
done
ls -la gpurun_out/t_*.ncu-rep
