#!/bin/bash
# first GPU pass of round 2: full -m gpu suite, bench (default + SWP issuer variant), op / KNN probes
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi -L >> gpurun_out/r2a_host.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
PU_LIB=libpointunet_b200_swp.so timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -q > gpurun_out/r2a_swp_pytest.log 2>&1; tail -2 gpurun_out/r2a_swp_pytest.log
PU_LIB=libpointunet_b200_swp.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_swp.json 2> gpurun_out/r2a_bench_swp.err; echo "bench swp rc=$?"
timeout 300 python tools/op_bench.py > gpurun_out/r2a_op_bench.jsonl 2>&1
timeout 300 python tools/knn_probe.py > gpurun_out/r2a_knn_probe.txt 2>&1
python -c "
import json
for f in ('r2a_bench.json','r2a_bench_swp.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('knn'))
    except Exception as e: print(f, 'ERR', e)
"
