#!/bin/bash
# Bottleneck analysis of K1 at 8192^3 (results with flags 2..16 are numerically wrong by design).
OUT=gpurun_out; mkdir -p $OUT
export EXPLORE_LOG=ablate.jsonl
for f in 0 1 2 4 8 16 17 9 10 12 24 28 30; do
  UGEMM_K1_FLAGS=$f timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 | sed "s/^/flags=$f /"
done
for kc in 1 2 4 8; do
  UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 $kc 0 8192 8192 8192 | sed "s/^/coll kc=$kc /"
done
UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 T N | sed "s/^/coll TN /"
UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 N T | sed "s/^/coll NT /"
UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 T T | sed "s/^/coll TT /"
UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 200704 256 1152 | sed "s/^/coll c4 /"
UGEMM_K1_FLAGS=1 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 16384 16384 16384 | sed "s/^/coll 16k /"
for t in "N N" "T N" "N T" "T T"; do UGEMM_K1_FLAGS=1 timeout 200 python tools/gpu_explore.py k1 2 $t | grep -c '"ok": true' | sed "s/^/coll correctness $t ok-count: /"; done
UGEMM_K1_FLAGS=1 timeout 200 python tools/gpu_explore.py k1 1 N N | grep -c '"ok": true' | sed "s/^/coll correctness cg1 NN ok-count: /"
