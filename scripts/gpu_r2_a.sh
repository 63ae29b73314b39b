# round 2, call A: gated tests + full-shape parity + parity table + reference eager probe + stream A/Bs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(XVA_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_full_gpu.py 2>&1 | tail -30) > gpurun_out/r2a_tests.log
tail -3 gpurun_out/r2a_tests.log
(XVA_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_parity_full_gpu.py -m gpu -q 2>&1 | tail -60) > gpurun_out/r2a_tests_full.log
tail -5 gpurun_out/r2a_tests_full.log
timeout 900 python scripts/parity_table.py > gpurun_out/r02_parity_table.txt 2> gpurun_out/r2a_parity_err.log
tail -5 gpurun_out/r02_parity_table.txt
timeout 600 python baseline/ref_step.py > gpurun_out/r2a_ref_probe.log 2>&1
tail -c 600 gpurun_out/r2a_ref_probe.log
for f in 0 1; do
  XVA_BWD_STREAMS=$f timeout 200 python bench.py --no-hifigan --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2a_streams_bench_$f.log 2>&1
  XVA_BWD_STREAMS=$f timeout 200 python bench.py --no-hifigan --no-cpu-baseline --no-graph --steps 30 --warmup 5 > gpurun_out/r2a_streams_bench_eager_$f.log 2>&1
done
python - <<'PY'
import json
for tag in ("0", "1", "eager_0", "eager_1"):
    try:
        d = json.loads(open(f"gpurun_out/r2a_streams_bench_{tag}.log").read().strip().splitlines()[-1])
        print("XVA_BWD_STREAMS", tag, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "frames/s")
    except Exception as e:
        print(tag, "failed", e)
PY
run_h() { tag=$1; shift; env "$@" timeout 300 python scripts/bench_hifigan.py 16 10 > gpurun_out/r2a_hifi_$tag.log 2>&1; echo "$tag: $(tail -1 gpurun_out/r2a_hifi_$tag.log | cut -c1-200)"; }
run_h base XVA_BWD_STREAMS=0
run_h side XVA_BWD_STREAMS=1
run_h disc2 XVA_DISC_STREAMS=2
run_h disc4 XVA_DISC_STREAMS=4
run_h disc8 XVA_DISC_STREAMS=8
run_h gen XVA_GEN_STREAMS=1
run_h all XVA_GEN_STREAMS=1 XVA_DISC_STREAMS=4 XVA_BWD_STREAMS=1
run_h base_eager XVA_NO_GRAPH=1
run_h all_eager XVA_NO_GRAPH=1 XVA_GEN_STREAMS=1 XVA_DISC_STREAMS=4 XVA_BWD_STREAMS=1
