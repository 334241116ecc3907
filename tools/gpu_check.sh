#!/bin/bash
# Staged GPU check of the tile engine: legacy stores -> TMA-store epilogue + fused stats -> CTA pairs -> full suite.
# Usage (on the GPU box): bash tools/gpu_check.sh [tag]
tag=${1:-chk}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; ( "$@" ) > gpurun_out/${tag}_$name.log 2>&1; echo "rc=$? $name"; tail -3 gpurun_out/${tag}_$name.log; }
run legacy env RG_CG2=0 RG_TMA_STORE=0 RG_FUSED_STATS=0 timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -k "not fused_stats"
run tma env RG_CG2=0 timeout 600 python -m pytest tests/test_engine_gpu.py -x -q
run cg2 timeout 600 python -m pytest tests/test_engine_gpu.py -x -q
run full timeout 1200 python -m pytest tests -x -q -m gpu
run mb_cg1 env RG_CG2=0 timeout 300 python tools/gemm_bench.py conv
run mb_cg2 timeout 300 python tools/gemm_bench.py
run bench timeout 900 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline
