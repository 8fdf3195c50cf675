#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -q --timeout=300 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
timeout 300 python tools/wgrad_bench.py > gpurun_out/r2e_wgrad_ts.jsonl 2>&1; tail -1 gpurun_out/r2e_wgrad_ts.jsonl
PU_WGRAD_TS=0 timeout 300 python tools/wgrad_bench.py > gpurun_out/r2e_wgrad_ss.jsonl 2>&1; tail -1 gpurun_out/r2e_wgrad_ss.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
PU_LIB=libpointunet_b200_swp.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_swp.json 2> gpurun_out/r2e_bench_swp.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('r2e_bench.json','r2e_bench_swp.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        b=d['breakdown_ms_per_step']
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
        print({k:v['ms'] for k,v in b.items()})
    except Exception as e: print(f, 'ERR', e)
PY
grep -h "L1 att1fc\|L2 att1fc\|L1 LFAmlp2\|fc1\|L3 att1fc" gpurun_out/r2e_wgrad_ts.jsonl
