#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/wgrad_bench.py > gpurun_out/r2d_wgrad_ts.jsonl 2>&1
PU_WGRAD_TS=0 timeout 300 python tools/wgrad_bench.py > gpurun_out/r2d_wgrad_ss.jsonl 2>&1
paste -d'|' <(cut -c1-130 gpurun_out/r2d_wgrad_ts.jsonl) <(python -c "
import json
for l in open('gpurun_out/r2d_wgrad_ss.jsonl'):
    d=json.loads(l); print(d.get('ms', d.get('total_ms')))")
