#!/bin/bash
# round 2, call AH: compute-sanitizer over the kernels added late in round 2 (vits.cu, sn_pack, float4 wn_pack, tiled conv_c1)
mkdir -p gpurun_out
SAN="compute-sanitizer --error-exitcode 9"
( timeout 500 $SAN --tool memcheck python -m pytest tests/test_vits_gpu.py -m gpu -q -x -k "gated or mel or wn_matches or flow or alignment" 2>&1 | tail -6 ) > gpurun_out/r2ah_memcheck_vits.log; echo "memcheck vits rc=$?"; tail -3 gpurun_out/r2ah_memcheck_vits.log | cut -c1-200
( timeout 500 $SAN --tool racecheck python -m pytest tests/test_hifigan_gpu.py tests/test_vits_gpu.py -m gpu -q -x -k "spectral_packer or packer or discriminator_forward or gated or alignment" 2>&1 | tail -6 ) > gpurun_out/r2ah_racecheck.log; echo "racecheck rc=$?"; tail -3 gpurun_out/r2ah_racecheck.log | cut -c1-200
