#!/bin/bash
# same-box per-kernel comparison, side stream off so that kernel durations are additive
export PYTHONUNBUFFERED=1 B200_WGRAD_STREAM=0
echo "== old"; (cd _old && timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair|attn_|ln_|colsum" | head -14)
echo "== new"; timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair|attn_|ln_|colsum|last micro" | head -16
