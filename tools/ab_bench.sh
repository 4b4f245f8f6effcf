export PYTHONUNBUFFERED=1
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json")); print("$name", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$name failed", e)
PY
}
run new A=1
run box2d B200_GEMM_LAB=32
run oldroles B200_GEMM_LAB=256
run nomask B200_DROP_MASK=0
run new2 A=1
run nographs B200_GRAPH=0
run nosidestream B200_WGRAD_STREAM=0
if [ -d _old ]; then
  (cd _old && timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > ../gpurun_out/ab_old.json 2> ../gpurun_out/ab_old.err)
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_old.json")); print("old-commit", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("old failed", e)
PY
fi
