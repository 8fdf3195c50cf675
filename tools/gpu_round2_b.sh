#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 300 python tools/knn_probe.py > gpurun_out/r2b_knn_probe.txt 2>&1
PU_KNN_DEFER=1 timeout 300 python tools/knn_probe.py > gpurun_out/r2b_knn_probe_defer.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:knn_search_kernel -c 1 -s 1 -f -o gpurun_out/r2b_knn_direct python tools/knn_one.py > gpurun_out/r2b_ncu1.log 2>&1
PU_KNN_DEFER=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:knn_search_kernel -c 1 -s 1 -f -o gpurun_out/r2b_knn_defer python tools/knn_one.py > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
head -12 gpurun_out/r2b_knn_probe.txt; head -12 gpurun_out/r2b_knn_probe_defer.txt
