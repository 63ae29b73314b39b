mkdir -p gpurun_out
pick() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'fastpitch', round(d['ms_per_step'],3), 'ms', round(d['roofline']['achieved'],1), 'TF/s')
PY
}
B="--steps 30 --warmup 5 --no-cpu-baseline --no-hifigan"
XVA_GEMM_HALO=0 timeout 300 python bench.py $B > gpurun_out/p_nohalo.log 2>&1; pick gpurun_out/p_nohalo.log no_halo
XVA_GEMM_HALO_ALIGN=8 timeout 300 python bench.py $B > gpurun_out/p_halo8.log 2>&1; pick gpurun_out/p_halo8.log halo_align8
timeout 300 python bench.py $B > gpurun_out/p_halo4.log 2>&1; pick gpurun_out/p_halo4.log halo_align4
