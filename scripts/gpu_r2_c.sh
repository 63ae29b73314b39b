# round 2, call C: stream-K schedule -- correctness, then A/B on the bench (one box, back to back)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -40) > gpurun_out/r2c_gemm_tests.log
tail -5 gpurun_out/r2c_gemm_tests.log
if grep -q "failed\|error" gpurun_out/r2c_gemm_tests.log; then echo "GEMM TESTS FAILED"; fi
b() { tag=$1; shift; env "$@" XVA_BENCH_GEMM_TABLE=gpurun_out/r2c_table_$tag.txt timeout 300 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2c_bench_$tag.log 2>&1; python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c_bench_{tag}.log").read().strip().splitlines()[-1])
    print(tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s  gemm", round(d["roofline"]["kernel_ms_per_step"], 3), "ms frac", round(d["roofline"]["frac"], 4), "loss", d["loss"])
except Exception as e:
    print(tag, "failed", e); print(open(f"gpurun_out/r2c_bench_{tag}.log").read()[-1500:])
PY
}
b sk0 XVA_GEMM_SK=0
b sk1 XVA_GEMM_SK=1
b sk1_min16 XVA_GEMM_SK=1 XVA_GEMM_SK_MIN=16
b sk1_min4 XVA_GEMM_SK=1 XVA_GEMM_SK_MIN=4
b sk1_ovh0 XVA_GEMM_SK=1 XVA_GEMM_TILE_OVH=0
b sk1_ovh16k XVA_GEMM_SK=1 XVA_GEMM_TILE_OVH=16000
(timeout 900 python -m pytest tests/test_fastpitch_gpu.py tests/test_parity_full_gpu.py tests/test_hifigan_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r2c_model_tests.log
tail -4 gpurun_out/r2c_model_tests.log
timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2c_hifi_sk1.log 2>&1; tail -1 gpurun_out/r2c_hifi_sk1.log | cut -c1-200
XVA_GEMM_SK=0 timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2c_hifi_sk0.log 2>&1; tail -1 gpurun_out/r2c_hifi_sk0.log | cut -c1-200
