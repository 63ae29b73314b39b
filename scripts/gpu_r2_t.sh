#!/bin/bash
# round 2, call T (2 GPUs): flow parity, hifi_only on 2 NCCL ranks, smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vits_gpu.py -q -m gpu -k "flow" 2>&1 | tail -25 > gpurun_out/r2t_flow.log; cut -c1-1500 gpurun_out/r2t_flow.log
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -q -m gpu -k "vits" 2>&1 | tail -25 > gpurun_out/r2t_ddp.log; cut -c1-1500 gpurun_out/r2t_ddp.log
cat gpurun_out/ddp_nccl_vits.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-800
