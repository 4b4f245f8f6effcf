#!/bin/bash
export PYTHONUNBUFFERED=1 B200_WGRAD_STREAM=0
echo "== new (plain launches)"; timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5
echo "== new, no pdl instructions (LAB=512)"; B200_GEMM_LAB=512 timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5
echo "== nolab build"; (cd _nolab && timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5)
echo "== old"; (cd _old && timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -4)
