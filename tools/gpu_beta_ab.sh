#!/bin/bash
# (record of an experiment: bits 21 / 23 and UGEMM_K1_CEV0 existed between commits 1ba5b77 and the next one only; results in profiles/r4b_ / r4e_beta_ab.jsonl)
# beta != 0 path of the TS kernel: 256-bit C preload (bit 20 = old 128-bit form) and TMA L2 prefetch of the next tile's C
# (bit 21 = old per-thread prefetch), one process per flag set, each set twice in alternating order.  usage: bash tools/gpu_beta_ab.sh <tag>
TAG=${1:-beta}
OUT=gpurun_out/${TAG}_beta_ab.jsonl; : > $OUT
for F in 1 3145729 1 3145729 1048577 2097153; do
  UGEMM_K1_FLAGS=$F timeout 60 python tools/gpu_beta_cases.py 2>&1 | tail -1 >> $OUT
done
cat $OUT
