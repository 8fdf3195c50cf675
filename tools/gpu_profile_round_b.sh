#!/bin/bash
# Second part of the round profile: --set full captures of the persistent tcgen05 kernel instances (selected by their
# demangled template arguments) and the per-op / per-shape tables.   bash tools/gpu_profile_round_b.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
full() {  # name, regex on the demangled kernel name
    timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
        -k "regex:$2" -c 2 -f -o $out/${tag}_full_$1 python tools/profile_step.py > $out/${tag}_full_$1.log 2>&1
    ncu -i $out/${tag}_full_$1.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$1.csv 2>/dev/null
}
full tc_persist_tma 'tc_persist_kernel<64, 0, '
full tc_att_bwd_fused 'tc_persist_kernel<64, 3, '
full tc_persist_stream 'tc_persist_kernel<128, 0, '
timeout 300 python tools/op_bench.py > $out/${tag}_op_bench.jsonl 2> $out/${tag}_op_bench.err
timeout 300 python tools/knn_probe.py > $out/${tag}_knn_probe.txt 2>&1
timeout 300 python tools/linear_bench.py > $out/${tag}_linear_shapes.jsonl 2> $out/${tag}_linear.err
timeout 300 python tools/wgrad_bench.py > $out/${tag}_wgrad_shapes.jsonl 2> $out/${tag}_wgrad.err
ls -la $out | grep ${tag}_
