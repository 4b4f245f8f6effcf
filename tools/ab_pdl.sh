#!/bin/bash
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abp_$1.json")); print("$1", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e: print("$1 failed", e)
PY
}
run() { name=$1; dir=$2; shift 2; (cd $dir && env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > /root/repo/gpurun_out/abp_$name.json 2> /root/repo/gpurun_out/abp_$name.err); show $name; }
run base . A=1
run pdl_all . B200_PDL=1
run pdl_gemm_only _lab B200_PDL=1
run base2 . A=1
