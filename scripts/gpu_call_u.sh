# session-5 GPU call U: CTC as three launches (alpha || beta), MAS log pre-pass + prefetch: parity, smoke, stage-1 timing
mkdir -p gpurun_out
(timeout 420 python -m pytest tests/test_stage1_gpu.py tests/test_mas_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/u_stage1.log
tail -4 gpurun_out/u_stage1.log
(timeout 200 python -m pytest tests/test_trainers_gpu.py -m gpu -q -k aligner 2>&1 | tail -30) > gpurun_out/u_trainer.log
tail -2 gpurun_out/u_trainer.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; tail -1 gpurun_out/u_smoke.log | cut -c1-400
timeout 300 python scripts/bench_stage1.py 20 > gpurun_out/u_stage1_bench.log 2>&1; tail -1 gpurun_out/u_stage1_bench.log | cut -c1-2200
