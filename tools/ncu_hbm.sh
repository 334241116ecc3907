#!/bin/bash
# ncu --set full capture of the HBM-bound kernels at the largest config-2 shapes; leaves a small CSV + summary in gpurun_out/
#   gpurun --timeout 900 -- 'bash tools/ncu_hbm.sh r2'
TAG=${1:-r2}
mkdir -p gpurun_out
REP=/tmp/hbm_kernels_$TAG
ncu --set full --clock-control none -k regex:"colreduce_stage1|ew_split|im2col_img|col2im_img|img_conv|adam_kernel" -c 40 \
    -o $REP python tools/hbm_bench.py --once --big > gpurun_out/${TAG}_ncu_hbm.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_hbm.log
python tools/ncu_summary.py $REP.ncu-rep > gpurun_out/${TAG}_ncu_hbm_summary.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_hbm_raw.csv 2>/dev/null
ls -la $REP.ncu-rep gpurun_out/${TAG}_ncu_hbm_raw.csv
sz=$(stat -c %s $REP.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp $REP.ncu-rep gpurun_out/${TAG}_hbm_kernels.ncu-rep; fi
cat gpurun_out/${TAG}_ncu_hbm_summary.txt
