# round 2, call I (2 GPUs): NCCL data-parallel parity test, bench at N = 2 (graph incl. NCCL), trainer facade under torchrun
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r2i_ddp_tests.log 2>&1
grep -E "ddp results|passed|failed|Error|assert" gpurun_out/r2i_ddp_tests.log | cut -c1-900 | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_n2.log 2> gpurun_out/r2i_bench_n2.err
echo "bench N=2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --via-trainer --steps 30 --warmup 5 > gpurun_out/r2i_bench_n2_via_trainer.log 2> gpurun_out/r2i_bench_n2_via_trainer.err
echo "via-trainer N=2 rc=$?"
python - <<'PY'
import json
for tag in ("bench_n2", "bench_n2_via_trainer"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2i_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        h = d.get("hifigan") or {}
        print(tag, d["n_gpus"], "GPUs", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s | hifigan", round(h.get("ms_per_step", 0), 2), "ms", round(h.get("value", 0)), d.get("config", {}).get("launch"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2i_{tag}.log").read()[-1500:]); print(open(f"gpurun_out/r2i_{tag}.err").read()[-1500:])
PY
