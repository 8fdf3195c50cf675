#!/bin/bash
# Cheap refresh of the round's numbers after a change that leaves the kernels of the --set full captures alone:
# launch list of one step, KNN probe, N = 1 bench line.   bash tools/gpu_profile_refresh.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
bash tools/gpu_profile_launchlist.sh > /dev/null 2>&1
for f in step_launches.csv step_launches.csv.gz summary_head.txt profile_step.log; do [ -f $out/r2c_$f ] && mv $out/r2c_$f $out/${tag}_$f; done
timeout 300 python tools/knn_probe.py > $out/${tag}_knn_probe.txt 2>&1
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
ls -la $out | grep ${tag}_ | head -20
