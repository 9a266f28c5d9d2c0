#!/bin/bash
# sweep of tools/mma_rate (one process per configuration); usage under gpurun: bash tools/gpu_mma_rate.sh <tag>
OUT=gpurun_out/${1:-r2}_mma_rate.jsonl; : > $OUT
for noise in 0 1 2; do for ts in 0 1; do for cg in 1 2; do for N in 64 128 192 256; do
  timeout 30 ./tools/mma_rate 4096 $ts $cg $N $noise 3 >> $OUT 2>&1 || echo "{\"failed\": [$ts, $cg, $N, $noise]}" >> $OUT
done; done; done; done
for ts in 0 1; do for cg in 1 2; do for N in 128 256; do timeout 30 ./tools/mma_rate 4096 $ts $cg $N 0 1 >> $OUT 2>&1; done; done; done
cat $OUT
