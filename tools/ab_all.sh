#!/bin/bash
# needs: git worktree add -f _old <commit> && (cd _old && python -m bert_pytorch_b200.ops.build)   [ab_all.sh also: a _lab/
# worktree built with B200_NVCC_EXTRA=-DB200_GEMM_LAB]; both directories are scratch (remove them afterwards)
# tests of the GEMM, then same-box: old commit vs this tree (bench + additive per-kernel trace), then lab cycles
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layer_grads.py tests/test_gpu_kernels.py -q -x 2>&1 | tail -4
show() { python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1.json")); print("$1", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$1 failed", e)
PY
}
(cd _old && timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > ../gpurun_out/ab_old.json 2> ../gpurun_out/ab_old.err); show old
timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err; show new
B200_WGRAD_WIDE=1 timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > gpurun_out/ab_wide.json 2> gpurun_out/ab_wide.err; show wide
echo "== new, additive trace"; B200_WGRAD_STREAM=0 timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair|last micro" | head -8
echo "== old, additive trace"; (cd _old && B200_WGRAD_STREAM=0 timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -4)
exit 0
echo "== lab"; (cd _lab && timeout 200 python tools/gemm_lab.py --wgrad --variants base,nomma,noepi 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['name'], {k:(v['cyc_per_kblock'] if isinstance(v,dict) else v) for k,v in d.items() if k not in ('name','M','N','K')})
")
