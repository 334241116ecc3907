#!/bin/bash
# run each engine test group in its own process so one trap does not mask the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/engine_probe.log 2>&1
for k in test_gemm_nt test_conv_down test_conv_up test_conv_up_img test_conv_wgrad test_proj_wgrad_and_fwd test_gemm_tn test_pack_edge; do
  echo "=== $k" >> gpurun_out/engine_probe.log
  timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "$k" --tb=short -p no:cacheprovider >> gpurun_out/engine_probe.log 2>&1
done
tail -c 6000 gpurun_out/engine_probe.log
