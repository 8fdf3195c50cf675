out=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:knn_query_warp_kernel" -c 1 -f -o $out/r2_full_knn_query_warp python tools/profile_step.py --mode knn > $out/r2_full_knn_query_warp.log 2>&1
ncu -i $out/r2_full_knn_query_warp.ncu-rep --page raw --csv > $out/r2_ncu_full_knn_query_warp.csv 2>/dev/null
ncu -i $out/r2_full_knn_query_warp.ncu-rep --page source --csv > $out/r2_ncu_src_knn_query_warp.csv 2>/dev/null
rm -f $out/r2_full_knn_query_warp.ncu-rep
ls -la $out | grep knn_query
