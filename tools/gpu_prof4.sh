#!/bin/bash
# per-role cycle breakdown of the TS kernel (flags 32768 + 32 + 1) and the SS kernel (33), a few shapes; clocks sampled alongside
TAG=${1:-r2i}
export EXPLORE_LOG=${TAG}_prof.jsonl UGEMM_K1_ABLATION=1
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 50 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
for f in 32801 33; do
  for shape in "8192 8192 8192" "4096 3072 2048" "4096 4096 4096"; do
    echo "=== flags=$f $shape"
    UGEMM_K1_FLAGS=$f timeout 120 python tools/gpu_explore.py time 3xtf32 2 4 0 $shape 2>&1 | grep -E "k1prof cta[0] |tflops" | tail -2 | cut -c1-520
  done
done
kill $SMI
sort -n gpurun_out/${TAG}_clocks.csv | uniq -c | sort -k1 -n -r | head -12
