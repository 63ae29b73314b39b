# session-4 GPU call G (2 GPUs): the bench line at N=2 (FastPitch + HiFi-GAN halves, gradient all-reduce paths)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g_bench_n2.log 2>&1
tail -1 gpurun_out/g_bench_n2.log | cut -c1-600
grep -E "Error|error|Traceback" gpurun_out/g_bench_n2.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/g_bench_ref_n2.log 2>&1
tail -1 gpurun_out/g_bench_ref_n2.log | cut -c1-300
