#!/bin/bash
# GPU check of the fused image-side kernels: parity tests, then the per-launch microbench.  bash tools/img_check.sh [tag]
tag=${1:-img}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_img_gpu.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
rc=$?; echo "tests rc=$rc"; tail -15 gpurun_out/${tag}_tests.log
if [ $rc -ne 0 ]; then
  echo "== retry with the cp.async tile loader"
  RG_IMG_TMA=0 timeout 300 python -m pytest tests/test_img_gpu.py -q -m gpu > gpurun_out/${tag}_tests_notma.log 2>&1
  echo "rc=$?"; tail -15 gpurun_out/${tag}_tests_notma.log
fi
timeout 200 python tools/hbm_bench.py img_conv > gpurun_out/${tag}_bench.txt 2>&1; cat gpurun_out/${tag}_bench.txt
timeout 200 python tools/synth_profile.py 2>&1 | grep -v -i warn | head -6
