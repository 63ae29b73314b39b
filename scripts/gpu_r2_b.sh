# round 2, call B (2 GPUs): whole GPU suite with nothing gated, NCCL data-parallel test, bench N=1 (eager_b200, reference arm), N=2 graph vs eager
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.log 2> gpurun_out/r2b_bench_n1.err
tail -c 1500 gpurun_out/r2b_bench_n1.log; tail -3 gpurun_out/r2b_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.log 2> gpurun_out/r2b_bench_ref.err
tail -c 600 gpurun_out/r2b_bench_ref.log
for g in 1 0; do
  XVA_BENCH_GRAPH_NCCL=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_n2_graph$g.log 2> gpurun_out/r2b_bench_n2_graph$g.err
  echo "N=2 graph=$g rc=$?"; tail -c 700 gpurun_out/r2b_bench_n2_graph$g.log; tail -5 gpurun_out/r2b_bench_n2_graph$g.err
done
