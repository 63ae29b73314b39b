# session-4 GPU call O: halo mode -- tests first (bounded), then A/B against XVA_GEMM_HALO=0 on the same box
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/o_gemm_tests.log
tail -4 gpurun_out/o_gemm_tests.log
if grep -q "passed" gpurun_out/o_gemm_tests.log && ! grep -q "failed" gpurun_out/o_gemm_tests.log; then
  (timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/o_tests.log
  tail -3 gpurun_out/o_tests.log
  pick() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
h=d.get('hifigan') or {}
print(sys.argv[2], 'fastpitch', round(d['ms_per_step'],3), 'ms', round(d['roofline']['achieved'],1), 'TF/s | hifigan', round(h.get('ms_per_step',0),2), 'ms', round(h.get('roofline',{}).get('achieved',0),1), 'TF/s')
PY
  }
  B="--steps 30 --warmup 5 --no-cpu-baseline"
  XVA_GEMM_HALO=0 timeout 400 python bench.py $B > gpurun_out/o_bench_nohalo.log 2>&1; pick gpurun_out/o_bench_nohalo.log no_halo
  XVA_BENCH_GEMM_TABLE=gpurun_out/o_fp_gemm_table.txt timeout 400 python bench.py $B > gpurun_out/o_bench_halo.log 2>&1; pick gpurun_out/o_bench_halo.log halo
  timeout 600 python scripts/bench_generator_large.py 8 880 gpurun_out/o_generator_large_table.txt > gpurun_out/o_gen_large.log 2>&1
  head -3 gpurun_out/o_generator_large_table.txt | cut -c1-200
fi
