#!/bin/bash
# round 2: per-role cycle breakdown (flags 32) + ablations + ncu full capture of the current build at 8192^3
TAG=${1:-r2d}
export EXPLORE_LOG=${TAG}_prof.jsonl UGEMM_K1_ABLATION=1
for f in 33 35 37 41 97; do
  echo "=== flags=$f"
  UGEMM_K1_FLAGS=$f timeout 120 python tools/gpu_explore.py time 3xtf32 2 4 0 8192 8192 8192 2>&1 | grep -E "k1prof cta[0] |tflops" | tail -2 | cut -c1-520
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k1ts_kernel|k1_3xtf32" -s 4 -c 1 -f -o gpurun_out/${TAG}_k1_c2 python bench.py --steps 5 --warmup 3 --no-shapes --no-ncu > gpurun_out/${TAG}_ncu_full_stdout.log 2>&1; echo "ncu full rc=$?"
