#!/bin/bash
# Round-end evidence run on the GPU box: bench line, ncu launch list of the same command, ncu --set full of the
# tile-engine kernels.  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 330 -c 24 -f -o gpurun_out/prof_gemm_${tag} \
  python bench.py --steps 1 --warmup 3 --e2e-steps 3 --synth-chunk 0 --vae-steps 0 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 1 --warmup 3 --e2e-steps 3 --synth-chunk 0 --vae-steps 0 --no-cpu-baseline > gpurun_out/ncu_list_${tag}.log 2>&1
echo "ncu list rc=$?"
