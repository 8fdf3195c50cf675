#!/bin/bash
# tests + micro-benchmark + ncu --set full of the fused position-branch kernels at level 0:  bash tools/gpu_locse_round.sh <tag>
tag=${1:-locse}
out=gpurun_out
mkdir -p $out
(timeout 300 python -m pytest tests/test_locse_mlp_gpu.py -x -q) > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
timeout 300 python tools/locse_bench.py 2>&1 | tee $out/${tag}_bench.jsonl
for k in ${NCU_KERNELS:-locse_mlp_fwd_kernel locse_mlp_bwd_kernel}; do
  LOCSE_LEVELS=${NCU_LEVEL:-0} timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $out/${tag}_full_$k \
      python tools/locse_bench.py > $out/${tag}_full_$k.log 2>&1
  ncu -i $out/${tag}_full_$k.ncu-rep --page raw --csv > $out/${tag}_ncu_$k.csv 2>/dev/null
done
