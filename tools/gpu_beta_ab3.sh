#!/bin/bash
# (record of an experiment: bits 21 / 23 and UGEMM_K1_CEV0 existed between commits 1ba5b77 and the next one only; results in profiles/r4b_ / r4e_beta_ab.jsonl)
# beta != 0 through the TMA unit: where in the tile the old C is added (UGEMM_K1_CEV0) and whether the L2 prefetch still pays (bit 23)
TAG=${1:-beta3}
OUT=gpurun_out/${TAG}_beta_ab.jsonl; : > $OUT
run() { echo -n "{\"cev0\": \"$1\", \"run\": " >> $OUT; UGEMM_K1_CEV0=$1 UGEMM_K1_FLAGS=$2 timeout 60 python tools/gpu_beta_cases.py 2>&1 | tail -1 | tr -d '\n' >> $OUT; echo "}" >> $OUT; }
run 0 1; run 12 1; run 24 1; run 0 8388609; run 12 8388609; run 0 1
cat $OUT
