#!/bin/bash
# per-role cycle breakdown (UGEMM_K1_FLAGS bit 5) for a few configurations at 8192^3; 1 timed iteration each
export EXPLORE_LOG=prof.jsonl
for f in 32 33 36 40 44; do
  echo "=== flags=$f"
  UGEMM_K1_FLAGS=$f timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 2>&1 | grep -E "k1prof cta[01] |tflops" | tail -3 | cut -c1-420
done
echo "=== flags=32 kc=0"
UGEMM_K1_FLAGS=32 timeout 120 python tools/gpu_explore.py time 3xtf32 2 0 0 8192 8192 8192 2>&1 | grep -E "k1prof cta[01] |tflops" | tail -3 | cut -c1-420
echo "=== flags=32 cg=1"
UGEMM_K1_FLAGS=32 timeout 120 python tools/gpu_explore.py time 3xtf32 1 2 0 8192 8192 8192 2>&1 | grep -E "k1prof cta[01] |tflops" | tail -3 | cut -c1-420
