# session-5 GPU call S: first run of the stage-1 aligner kernels + tap-inner operand order A/B
mkdir -p gpurun_out
(timeout 420 python -m pytest tests/test_stage1_gpu.py -m gpu -q 2>&1 | tail -60) > gpurun_out/s_stage1.log
tail -5 gpurun_out/s_stage1.log
(timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_fastpitch_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/s_gemm_fp.log
tail -3 gpurun_out/s_gemm_fp.log
(timeout 200 python -m pytest tests/test_trainers_gpu.py -m gpu -q -k aligner 2>&1 | tail -30) > gpurun_out/s_trainer.log
tail -3 gpurun_out/s_trainer.log
XVA_GEMM_TAP_INNER=0 XVA_BENCH_GEMM_TABLE=gpurun_out/s_table_tap_outer.txt timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/s_bench_tap_outer.log 2>&1
XVA_GEMM_TAP_INNER=1 XVA_BENCH_GEMM_TABLE=gpurun_out/s_table_tap_inner.txt timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/s_bench_tap_inner.log 2>&1
python - <<'PY'
import json
for tag in ("tap_outer", "tap_inner"):
    try:
        d = json.loads(open(f"gpurun_out/s_bench_{tag}.log").read().strip().splitlines()[-1])
        dl = d["roofline"]["dominant_launch"]
        print(tag, round(d["ms_per_step"], 3), "ms; gemm", round(d["roofline"]["achieved"], 1), "TF/s; dominant", dl["shape"]["mode"], dl["shape"]["K"], round(dl["us_per_launch"], 1), "us", round(dl["frac"], 3))
    except Exception as e:
        print(tag, "failed", e)
PY
