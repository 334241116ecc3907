#!/bin/bash
# First GPU call for the B-resident tile-engine plan (RG_BRES=1), one GPU:
#   gpurun --timeout 420 -- 'bash tools/bres_ab.sh'
# 1. the gated bit-identity test, 2. the conv micro-benchmarks with and without the plan (RG_BRES=1: the strided L1
# layer; RG_BRES=2: also the merged-phase transposed L1 layer -- the only ones that qualify at the lung shapes),
# 3. A/B of the whole step.  Every leg runs under its own timeout: a barrier bug would hang.
mkdir -p gpurun_out
echo "== gated test"
RG_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k b_resident 2>&1 | tail -5
for f in 0 2; do
  echo "== conv micro-benchmark, RG_BRES=$f"
  RG_BRES=$f timeout 150 python tools/gemm_bench.py conv_ 2>&1 | grep -i "L1\|conv_" | head -14
done
S='import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],"ms/step",round(d["ms_per_step"],3),"frac",round(d["roofline"]["frac"],4))'
for f in 0 2 0 2; do
  RG_BRES=$f timeout 120 python bench.py --no-cpu-baseline --synth-chunk 0 --vae-steps 0 > "gpurun_out/bres_$f.json" 2> "gpurun_out/bres_$f.err" \
    && python -c "$S" "gpurun_out/bres_$f.json" || tail -5 "gpurun_out/bres_$f.err"
done
