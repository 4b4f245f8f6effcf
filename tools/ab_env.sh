#!/bin/bash
# same-box A/B of environment toggles on the headline micro-step: tools/ab_env.sh "NAME=VAL ..." "NAME2=VAL2" ...
export PYTHONUNBUFFERED=1
i=0
for envs in "A=1" "$@" "A=1"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > gpurun_out/abe_$i.json 2> gpurun_out/abe_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abe_$i.json")); print("$envs", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e: print("$envs failed", e)
PY
done
