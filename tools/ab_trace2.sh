#!/bin/bash
export PYTHONUNBUFFERED=1 B200_WGRAD_STREAM=0
for lab in 0 32 256 288; do
  echo "== B200_GEMM_LAB=$lab"; B200_GEMM_LAB=$lab timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5
done
echo "== PROMO=0"; B200_TMAP_PROMO=0 timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5
echo "== PROMO=2"; B200_TMAP_PROMO=2 timeout 300 python tools/trace_step.py 2>&1 | grep -E "phase 1|gemm_pair" | head -5
