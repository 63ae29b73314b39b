# round 2, final check on one GPU: the whole GPU suite, smoke(), the driver's bench command, the reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2final_gpu_tests.log 2>&1; tail -6 gpurun_out/r2final_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final_smoke.log 2>&1; tail -2 gpurun_out/r2final_smoke.log | cut -c1-300
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2final_bench_n1.log 2> gpurun_out/r2final_bench_n1.err; tail -3 gpurun_out/r2final_bench_n1.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2final_bench_ref.log 2> gpurun_out/r2final_bench_ref.err; tail -3 gpurun_out/r2final_bench_ref.err
python - <<'PY'
import json
for tag in ("bench_n1", "bench_ref"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2final_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        h = d.get("hifigan") or {}
        r = d.get("roofline") or {}
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "| hifigan", round(h.get("ms_per_step", 0), 2), "ms", round(h.get("value", 0)),
              "| frac", r.get("frac"), "attn", (r.get("attention") or {}).get("frac"), "combined", (r.get("tensor_kernels_combined") or {}).get("frac"),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), "| eager", {k: round(v["ms_per_step"], 1) for k, v in ((d.get("eager_b200") or {}).get("fastpitch") or {}).items()}, "clocks", d.get("clocks"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2final_{tag}.log").read()[-1500:])
PY
