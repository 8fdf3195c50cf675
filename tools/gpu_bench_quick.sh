#!/bin/bash
# quick A/B of the train step: bash tools/gpu_bench_quick.sh <tag> [ENV=VAL ...]
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"])
    print("  ", {k:v["ms"] for k,v in d["breakdown_ms_per_step"].items()})
except Exception as e:
    print(tag, "ERR", e); print(open(f"gpurun_out/{tag}.err").read()[-1500:])
PY
