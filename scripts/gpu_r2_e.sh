# round 2, call E: fused attention backward + integration
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -s > gpurun_out/r2e_attn_tests.log 2>&1
tail -30 gpurun_out/r2e_attn_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_fastpitch_gpu.py tests/test_parity_full_gpu.py tests/test_infer_gpu.py tests/test_stage1_gpu.py -m gpu -q > gpurun_out/r2e_fp_tests.log 2>&1
tail -15 gpurun_out/r2e_fp_tests.log | cut -c1-300
for f in 0 1; do
  XVA_FUSED_ATTN=$f XVA_BENCH_GEMM_TABLE=gpurun_out/r2e_table_attn$f.txt timeout 300 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2e_bench_attn$f.log 2>&1
done
python - <<'PY'
import json
for tag in ("attn0", "attn1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2e_bench_{tag}.log").read().splitlines() if l.startswith("{")][-1])
        print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), "gemm", round(d["roofline"]["kernel_ms_per_step"], 3), "ms frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"] // d["steps"], "loss", d["loss"])
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r2e_bench_{tag}.log").read()[-1500:])
PY
