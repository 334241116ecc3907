#!/bin/bash
# First GPU call for the copy-engine gradient exchange (parallel.PeerExchange, RG_DP_EXCHANGE=ce), N GPUs of one box:
#   gpurun --gpus 2 --timeout 420 -- 'bash tools/dp_exchange_ab.sh 2'
# 1. the gated GPU tests (rg_slices_sum vs an ordered sum; 2-GPU parity run through the exchange),
# 2. A/B of the bench at N ranks: NCCL all-reduce (default) vs the exchange, same box, short legs.
N=${1:-2}
mkdir -p gpurun_out
echo "== experimental tests"
RG_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -5
S='import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],"steps/s",round(d["value"],2),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],2))'
for mode in nccl ce nccl ce; do
  echo "== RG_DP_EXCHANGE=$mode, N=$N"
  RG_DP_EXCHANGE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus "$N" --no-cpu-baseline --synth-chunk 0 --vae-steps 0 \
    > "gpurun_out/dp_${mode}_n${N}.json" 2> "gpurun_out/dp_${mode}_n${N}.err" \
    && python -c "$S" "gpurun_out/dp_${mode}_n${N}.json" || tail -5 "gpurun_out/dp_${mode}_n${N}.err"
done
