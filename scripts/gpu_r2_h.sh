# round 2, call H: attention after the code-size fix, trainer facades with graph replay, bench through both paths
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -s -x > gpurun_out/r2h_attn_tests.log 2>&1
rc=$?
grep -E "attn_|passed|failed" gpurun_out/r2h_attn_tests.log | cut -c1-200
if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2h_attn_tests.log; echo "ATTENTION TESTS rc=$rc: stopping"; exit 0; fi
timeout 400 python -m pytest tests/test_trainers_gpu.py -m gpu -q > gpurun_out/r2h_trainer_tests.log 2>&1; tail -12 gpurun_out/r2h_trainer_tests.log | cut -c1-300
XVA_BENCH_GEMM_TABLE=gpurun_out/r2h_table.txt timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2h_bench.log 2>&1
timeout 300 python bench.py --via-trainer --steps 30 --warmup 5 > gpurun_out/r2h_bench_via_trainer.log 2>&1
XVA_TRAINER_GRAPH=0 timeout 300 python bench.py --via-trainer --steps 30 --warmup 5 > gpurun_out/r2h_bench_via_trainer_eager.log 2>&1
python - <<'PY'
import json
for tag in ("bench", "bench_via_trainer", "bench_via_trainer_eager"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2h_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        h = d.get("hifigan") or {}
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s | hifigan", round(h.get("ms_per_step", 0), 2), "ms", round(h.get("value", 0)),
              "| roofline", {k: round(v, 4) if isinstance(v, float) else v for k, v in (d.get("roofline") or {}).items() if k in ("frac", "achieved", "kernel_ms_per_step")},
              "attn", {k: round(v, 4) for k, v in ((d.get("roofline") or {}).get("attention") or {}).items() if k in ("frac", "achieved", "kernel_ms_per_step")},
              d.get("logged_frames_per_s_last_step"), d.get("host_wall_ms_per_step"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2h_{tag}.log").read()[-2000:])
PY
