#!/bin/bash
# lean issue loop (elect.sync + incremental descriptors) against the lane-0 loop; args: kblocks ts cg N noise nmma bmn accswap epi window lean
OUT=gpurun_out/${1:-r2}_mma_rate3.jsonl; : > $OUT
run() { timeout 30 ./tools/mma_rate "$@" >> $OUT 2>&1 || echo "{\"failed\": \"$*\"}" >> $OUT; }
for lean in 0 1; do
run 4096 0 2 256 0 3 1 1 0 2 $lean
run 4096 0 2 128 0 3 1 1 0 4 $lean
run 4096 1 2 128 0 3 1 1 0 4 $lean
run 4096 0 1 128 0 3 1 1 0 4 $lean
run 4096 1 1 128 0 3 1 1 0 4 $lean
run 4096 0 2 64 0 3 1 1 0 6 $lean
run 4096 1 2 64 0 3 1 1 0 6 $lean
run 4096 1 2 128 2 3 1 1 2 4 $lean
run 4096 0 2 192 2 3 1 0 2 4 $lean
run 4096 1 2 192 2 3 1 0 2 4 $lean
done
cat $OUT
