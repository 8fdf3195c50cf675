#!/bin/bash
# Round profile, run on the GPU box:  bash tools/gpu_profile_round.sh <round tag, e.g. r2>
# 1. ncu launch list of ONE training step (device time + DRAM bytes per launch)  -> profiles/<tag>_step_launches.csv.gz + summary
# 2. ncu --set full of the dominant kernels (a few launches each)                 -> gpurun_out/<tag>_full_*.ncu-rep (+ raw csv)
# Numbers printed under ncu are never bench values.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $out/${tag}_step_launches.csv python tools/profile_step.py > $out/${tag}_profile_step.log 2>&1
python tools/summarize_launches.py $out/${tag}_step_launches.csv ${tag}_step > $out/${tag}_summary_head.txt 2>&1
gzip -f -k $out/${tag}_step_launches.csv
full() {  # name, kernel regex, launches to skip
    timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" -s $3 -c 2 -f \
        -o $out/${tag}_full_$1 python tools/profile_step.py > $out/${tag}_full_$1.log 2>&1
    ncu -i $out/${tag}_full_$1.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$1.csv 2>/dev/null
}
full tc_persist_tma 'tc_persist_kernel<64, 0, false' 0
full tc_att_bwd_fused 'tc_persist_kernel<64, 3' 0
full tc_wgrad_ts 'tc_wgrad_ts_kernel' 0
full bn_bwd_apply 'bn_bwd_apply_kernel' 0
ls -la $out | grep ${tag}_
