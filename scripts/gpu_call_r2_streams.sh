# round-2 opener: the two-stream backward experiment (XVA_BWD_STREAMS=1, fastpitch.py::_wgrad_side) -- parity, then A/B
mkdir -p gpurun_out
(XVA_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_fastpitch_gpu.py -m gpu -q -k two_stream 2>&1 | tail -20) > gpurun_out/r2_streams_test.log
tail -3 gpurun_out/r2_streams_test.log
for f in 0 1; do
  XVA_BWD_STREAMS=$f timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2_streams_bench_$f.log 2>&1
  XVA_BWD_STREAMS=$f timeout 200 python bench.py --no-hifigan --no-cpu-baseline --no-graph --steps 30 --warmup 5 > gpurun_out/r2_streams_bench_eager_$f.log 2>&1
done
python - <<'PY'
import json
for tag in ("0", "1", "eager_0", "eager_1"):
    try:
        d = json.loads(open(f"gpurun_out/r2_streams_bench_{tag}.log").read().strip().splitlines()[-1])
        print("XVA_BWD_STREAMS", tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s")
    except Exception as e:
        print(tag, "failed", e)
PY
# the HiFi-GAN half (hifigan._Side): parity, then A/B of the replayed and the eager step
(XVA_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_hifigan_gpu.py -m gpu -q -k two_stream 2>&1 | tail -20) > gpurun_out/r2_streams_hifigan_test.log
tail -3 gpurun_out/r2_streams_hifigan_test.log
for f in 0 1; do
  XVA_BWD_STREAMS=$f timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_streams_hifigan_$f.log 2>&1; tail -1 gpurun_out/r2_streams_hifigan_$f.log | cut -c1-160
  XVA_BWD_STREAMS=$f XVA_NO_GRAPH=1 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_streams_hifigan_eager_$f.log 2>&1; tail -1 gpurun_out/r2_streams_hifigan_eager_$f.log | cut -c1-160
done
# sub-discriminators on parallel streams (hifigan._Branches), alone and with the side-stream weight gradients
for n in 2 4 8; do
  XVA_DISC_STREAMS=$n timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_disc_streams_$n.log 2>&1; tail -1 gpurun_out/r2_disc_streams_$n.log | cut -c1-160
done
XVA_DISC_STREAMS=4 XVA_BWD_STREAMS=1 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_disc_streams_4_side.log 2>&1; tail -1 gpurun_out/r2_disc_streams_4_side.log | cut -c1-160
XVA_GEN_STREAMS=1 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_gen_streams.log 2>&1; tail -1 gpurun_out/r2_gen_streams.log | cut -c1-160
XVA_GEN_STREAMS=1 XVA_DISC_STREAMS=4 XVA_BWD_STREAMS=1 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2_all_streams.log 2>&1; tail -1 gpurun_out/r2_all_streams.log | cut -c1-160
# multi-GPU (run with gpurun --gpus 2): NCCL all-reduces captured inside the step graph
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline
#   XVA_BENCH_GRAPH_NCCL=1 python -m torch.distributed.run ... (same line)
# warm-cache captures of the two sequential stage-1 kernels (the round-1 captures flushed the caches between replays)
for k in ctc_recursion mas; do
  timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o gpurun_out/r2_warm_$k python scripts/bench_stage1.py 1 --no-cpu > gpurun_out/r2_prof_warm_$k.log 2>&1
done
# voice-folder loader end to end (gated test) -- ungate it in tests/test_trainers_gpu.py once green
(XVA_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_trainers_gpu.py -m gpu -q -k voice_folder 2>&1 | tail -15) > gpurun_out/r2_voice_folder.log
tail -3 gpurun_out/r2_voice_folder.log
