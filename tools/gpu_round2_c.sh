#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_knn_gpu.py -q --timeout=300 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2c_bench_ts.json 2> gpurun_out/r2c_bench_ts.err; echo "bench rc=$?"
PU_WGRAD_TS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2c_bench_ss.json 2> gpurun_out/r2c_bench_ss.err; echo "bench rc=$?"
timeout 300 python tools/knn_probe.py > gpurun_out/r2c_knn_probe.txt 2>&1
python - <<'PY'
import json
for f in ('r2c_bench_ts.json','r2c_bench_ss.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        b=d['breakdown_ms_per_step']
        print(f, d['ms_per_step'], d['e2e']['ms_per_step'], 'wgrad', b.get('pu_tc_wgrad'), 'lin', b.get('pu_tc_linear_fwd'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2c_bench_ts.err
cat gpurun_out/r2c_knn_probe.txt
