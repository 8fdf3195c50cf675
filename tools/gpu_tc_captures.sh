tag=r2
out=gpurun_out
mkdir -p $out
full() {
    timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
        -k "regex:$2" -c 2 -f -o $out/${tag}_full_$1 python tools/profile_step.py > $out/${tag}_full_$1.log 2>&1
    ncu -i $out/${tag}_full_$1.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$1.csv 2>/dev/null
    rm -f $out/${tag}_full_$1.ncu-rep
}
full tc_persist_stream 'tc_persist_kernel<\(int\)128, \(int\)0, \(bool\)1'
full tc_persist_tma 'tc_persist_kernel<\(int\)64, \(int\)0, '
full tc_att_bwd_fused 'tc_persist_kernel<\(int\)64, \(int\)3, '
ls -la $out | grep ncu_full_tc
