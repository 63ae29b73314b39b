# round 2, final multi-GPU check: NCCL parity test (2 ranks) and the driver's bench command at N = world
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${XVA_N:-2}
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r2final_ddp_tests.log 2>&1
  grep -E "ddp results|passed|failed|Error|assert" gpurun_out/r2final_ddp_tests.log | cut -c1-700 | head -12
fi
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2final_bench_n$N.log 2> gpurun_out/r2final_bench_n$N.err
echo "bench N=$N rc=$?"; tail -4 gpurun_out/r2final_bench_n$N.err | cut -c1-200
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2final_bench_n{n}.log").read().splitlines() if l.startswith("{")][-1])
    h = d.get("hifigan") or {}
    print("N", d["n_gpus"], round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), "| hifigan", round(h.get("ms_per_step", 0), 2), "ms", round(h.get("value", 0)), round((h.get("e2e") or {}).get("value", 0)), d["config"]["launch"], d.get("clocks"))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/r2final_bench_n{n}.log").read()[-1500:])
PY
