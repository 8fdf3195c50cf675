#!/bin/bash
# Final profile of a round, run on the GPU box:  bash tools/gpu_profile_final.sh <tag>
#   launch list of one step (+ summary, traffic json), --set full captures of the kernels DESIGN.md discusses, the per-op /
#   per-shape tables, the KNN probe, and the two bench arms.  Numbers printed under ncu are never bench values.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
bash tools/gpu_profile_launchlist.sh > /dev/null 2>&1
for f in step_launches.csv step_launches.csv.gz summary_head.txt profile_step.log; do [ -f $out/r2c_$f ] && mv $out/r2c_$f $out/${tag}_$f; done
python tools/summarize_launches.py $out/${tag}_step_launches.csv ${tag}_step > $out/${tag}_summary_head.txt 2>&1
full() {  # name, regex on the demangled kernel name
    timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
        -k "regex:$2" -c 2 -f -o $out/${tag}_full_$1 python tools/profile_step.py > $out/${tag}_full_$1.log 2>&1
    ncu -i $out/${tag}_full_$1.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$1.csv 2>/dev/null
    rm -f $out/${tag}_full_$1.ncu-rep
}
full tc_persist_stream 'tc_persist_kernel<\(int\)128, \(int\)0, \(bool\)1'
full tc_persist_tma 'tc_persist_kernel<\(int\)64, \(int\)0, '
full tc_att_bwd_fused 'tc_persist_kernel<\(int\)64, \(int\)3, '
full locse_mlp_bwd 'locse_mlp_bwd_kernel'
full locse_mlp_fwd 'locse_mlp_fwd_kernel'
full bn_bwd_apply 'bn_bwd_apply_kernel'
timeout 300 python tools/op_bench.py > $out/${tag}_op_bench.jsonl 2> $out/${tag}_op_bench.err
timeout 300 python tools/locse_bench.py > $out/${tag}_locse_bench.jsonl 2> $out/${tag}_locse_bench.err
timeout 300 python tools/knn_probe.py > $out/${tag}_knn_probe.txt 2>&1
timeout 300 python tools/linear_bench.py > $out/${tag}_linear_shapes.jsonl 2> $out/${tag}_linear.err
timeout 300 python tools/wgrad_bench.py > $out/${tag}_wgrad_shapes.jsonl 2> $out/${tag}_wgrad.err
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
ls -la $out | grep ${tag}_
