mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_regulate_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t_regulate.log
for t in test_conv_fwd test_conv_fwd_relu_lens_strided test_conv_fwd_layernorm test_relu_then_layernorm test_conv_dgrad test_dgrad_relu_gate test_conv_wgrad test_attention_bmms test_dropout_matches test_full_size_linearity; do
  echo "=== $t" >> gpurun_out/t_gemm.log
  timeout 200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "$t" -s 2>&1 | grep -v "^$" | tail -60 >> gpurun_out/t_gemm.log
  echo "exit $?" >> gpurun_out/t_gemm.log
done
timeout 600 python baseline/_ref/probe_ref_eager.py > gpurun_out/probe.log 2>&1
tail -5 gpurun_out/t_regulate.log; grep -E "===|passed|failed|exit|Error|error" gpurun_out/t_gemm.log | head -60
