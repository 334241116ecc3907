#!/bin/bash
# Data-parallel A/B on N GPUs of one box:   gpurun --gpus N --timeout 900 -- 'bash tools/dp_exchange_ab.sh N'
# 1. (N == 2) the 2-GPU parity tests (NCCL and the copy-engine exchange),
# 2. bench legs: gradient exchange {nccl, ce} x optimiser-step deferral {1, 0} (steps.DEFER_UPDATES), short legs.
N=${1:-2}
LEGS=${2:-"nccl:1 ce:1 nccl:0 ce:0"}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  echo "== 2-GPU parity tests"
  timeout 400 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -5
fi
S='import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],"steps/s",round(d["value"],2),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],2),"sustained",round(d["sustained"]["value"],2) if d.get("sustained") else None,"dp_check",d.get("dp_check"))'
for leg in $LEGS; do
  mode=${leg%%:*}; defer=${leg##*:}
  echo "== RG_DP_EXCHANGE=$mode RG_DP_DEFER=$defer, N=$N"
  out="gpurun_out/dp_${mode}_d${defer}_n${N}"
  RG_DP_EXCHANGE=$mode RG_DP_DEFER=$defer timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
    --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus "$N" --no-cpu-baseline --no-stock --synth-chunk 0 \
    --vae-steps 0 --sustained-steps 100 > "$out.json" 2> "$out.err" \
    && python -c "$S" "$out.json" || tail -5 "$out.err"
done
