# session-4 GPU call L: north-star kernel size for the generator (880-frame mels), ragged / stage-4 / stage-2 FastPitch lines
mkdir -p gpurun_out
timeout 600 python scripts/bench_generator_large.py 8 880 gpurun_out/l_generator_large_table.txt > gpurun_out/l_gen_large.log 2>&1
head -40 gpurun_out/l_generator_large_table.txt | cut -c1-160
tail -3 gpurun_out/l_gen_large.log | cut -c1-300
B="--steps 20 --warmup 5 --no-cpu-baseline --no-hifigan"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'frames/s', round(d['roofline']['achieved'],1), 'TF/s gemm')" "$1" "$2" 2>&1 | tail -1; }
timeout 300 python bench.py $B --ragged > gpurun_out/l_bench_ragged.log 2>&1; pick gpurun_out/l_bench_ragged.log ragged
timeout 300 python bench.py $B --stage 4 > gpurun_out/l_bench_stage4.log 2>&1; pick gpurun_out/l_bench_stage4.log stage4
timeout 300 python bench.py $B --stage 2 > gpurun_out/l_bench_stage2.log 2>&1; pick gpurun_out/l_bench_stage2.log stage2
