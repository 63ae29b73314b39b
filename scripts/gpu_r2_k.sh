mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2k_small.txt
v() { tag=$1; shift; env "$@" timeout 120 python scripts/bench_gemm_small.py $tag >> gpurun_out/r2k_small.txt 2>> gpurun_out/r2k_small.err; }
v default
v nt64 XVA_GEMM_NTILE=64
v nt128 XVA_GEMM_NTILE=128
v nt256 XVA_GEMM_NTILE=256
v nt512 XVA_GEMM_NTILE=512
v pair2 XVA_GEMM_PAIR=2
v seg0 XVA_GEMM_SEG=0
v pdl XVA_GEMM_PDL=1
cat gpurun_out/r2k_small.txt; tail -3 gpurun_out/r2k_small.err
