# session-4 GPU call M: programmatic dependent launch + row-chunked weight-gradient split
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/m_tests.log
tail -3 gpurun_out/m_tests.log
pick() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
h=d.get('hifigan') or {}
print(sys.argv[2], 'fastpitch', round(d['ms_per_step'],3), 'ms', round(d['roofline']['achieved'],1), 'TF/s | hifigan', round(h.get('ms_per_step',0),2), 'ms', round(h.get('roofline',{}).get('achieved',0),1), 'TF/s')
PY
}
B="--steps 30 --warmup 5 --no-cpu-baseline"
XVA_GEMM_PDL=0 timeout 400 python bench.py $B > gpurun_out/m_bench_nopdl.log 2>&1; pick gpurun_out/m_bench_nopdl.log no_pdl
timeout 400 python bench.py $B > gpurun_out/m_bench_pdl.log 2>&1; pick gpurun_out/m_bench_pdl.log pdl
timeout 600 python scripts/bench_generator_large.py 8 880 gpurun_out/m_generator_large_table.txt > gpurun_out/m_gen_large.log 2>&1
head -12 gpurun_out/m_generator_large_table.txt | cut -c1-170
timeout 300 python scripts/prof_hifigan.py 16 gpurun_out/m_hifigan_gemm_table.txt > gpurun_out/m_hifigan_prof.log 2>&1
head -8 gpurun_out/m_hifigan_gemm_table.txt | cut -c1-120
