# round 2, call D: stream-K microbenchmarks (policy variants), fused attention forward tests, GEMM tests with full log
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_attn_fused_gpu.py -m gpu -q -s -x > gpurun_out/r2d_attn_tests.log 2>&1
tail -25 gpurun_out/r2d_attn_tests.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q > gpurun_out/r2d_gemm_tests.log 2>&1
tail -5 gpurun_out/r2d_gemm_tests.log
v() { tag=$1; shift; env "$@" timeout 200 python scripts/bench_gemm_sk.py $tag >> gpurun_out/r2d_sk_micro.txt 2>> gpurun_out/r2d_sk_micro.err; }
v default
v nocost XVA_GEMM_SK_COST=0
v nocost_split2 XVA_GEMM_SK_COST=0 XVA_GEMM_SK_MAXSPLIT=2
v nocost_split3 XVA_GEMM_SK_COST=0 XVA_GEMM_SK_MAXSPLIT=3
v nocost_min24 XVA_GEMM_SK_COST=0 XVA_GEMM_SK_MIN=24
v split8 XVA_GEMM_SK_MAXSPLIT=8
v nt384 XVA_GEMM_NTILE=512
v nt256 XVA_GEMM_NTILE=256
v nt128 XVA_GEMM_NTILE=128
cat gpurun_out/r2d_sk_micro.txt
tail -5 gpurun_out/r2d_sk_micro.err
