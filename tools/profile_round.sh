#!/bin/bash
# Round-end evidence run on the GPU box: bench line, ncu launch list of the same command, ncu --set full of the
# tile-engine kernels (summaries only travel back: gpurun_out is capped at 64 MiB).  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"
# the profiler runs use eager launches (RG_GRAPHS=0): same kernels, same order as the replayed graphs
ARGS="--steps 1 --warmup 3 --e2e-steps 3 --synth-chunk 0 --vae-steps 0 --no-cpu-baseline --no-stock --sustained-steps 0"
RG_GRAPHS=0 timeout 900 ncu --set full --clock-control none -k regex:gemm_ -s 310 -c 24 -f -o /tmp/prof_gemm_${tag} \
  python bench.py $ARGS > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/prof_gemm_${tag}.ncu-rep gpurun_out/ncu_traffic_${tag}.json > gpurun_out/ncu_gemm_${tag}.txt 2>&1
RG_GRAPHS=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_${tag}.csv \
  python bench.py $ARGS > gpurun_out/ncu_list_${tag}.log 2>&1
echo "ncu list rc=$?"
python tools/launch_summary.py /tmp/launches_${tag}.csv 11 > gpurun_out/launches_summary_${tag}.txt 2>&1
gzip -c /tmp/launches_${tag}.csv > gpurun_out/launches_${tag}.csv.gz
python tools/step_profile.py > gpurun_out/step_profile_${tag}.txt 2>&1
python tools/synth_profile.py > gpurun_out/synth_profile_${tag}.txt 2>&1
python tools/vae_profile.py > gpurun_out/vae_profile_${tag}.txt 2>&1
ls -la gpurun_out | tail -12
