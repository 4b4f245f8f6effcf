#!/bin/bash
# needs: git worktree add -f _old <commit> && (cd _old && python -m bert_pytorch_b200.ops.build)   [ab_all.sh also: a _lab/
# worktree built with B200_NVCC_EXTRA=-DB200_GEMM_LAB]; both directories are scratch (remove them afterwards)
# same-box A/B of the headline micro-step: current tree vs the tree checked out under _old/ (git worktree of an earlier commit)
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1.json")); print("$1", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
except Exception as e: print("$1 failed", e)
PY
}
for i in 1 2; do
  (cd _old && timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > ../gpurun_out/ab_old$i.json 2> ../gpurun_out/ab_old$i.err); show old$i
  timeout 200 python bench.py --steps 3 --warmup 3 --accum 8 --no-extras --no-e2e > gpurun_out/ab_new$i.json 2> gpurun_out/ab_new$i.err; show new$i
done
(cd _old && timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair|attn_|ln_" | head -12)
