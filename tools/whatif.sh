#!/bin/bash
# what-if experiments on the forward tile engine: which stage of the pipeline bounds a layer
for w in 0 1 2 4 8 3 12 6; do
  echo "== RG_WHATIF=$w (1 no MMA, 2 no epilogue, 4 no A loads, 8 no B loads)"
  RG_WHATIF=$w timeout 120 python tools/gemm_bench.py "$1" 2>&1 | grep -v wgrad
done
