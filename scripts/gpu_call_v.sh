# session-5 GPU call V: eight-step register look-ahead in the CTC recursion and MAS: parity, smoke, stage-1 timing
mkdir -p gpurun_out
(timeout 420 python -m pytest tests/test_stage1_gpu.py tests/test_mas_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/v_stage1.log
tail -4 gpurun_out/v_stage1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; tail -1 gpurun_out/v_smoke.log | cut -c1-400
timeout 300 python scripts/bench_stage1.py 20 > gpurun_out/v_stage1_bench.log 2>&1; tail -1 gpurun_out/v_stage1_bench.log | cut -c1-2200
