#!/bin/bash
# second sweep of tools/mma_rate: what in K1's MMA stream costs more than the 128-cycle floor of a 256x256x8 TF32 MMA?
# args: kblocks ts cg N noise nmma bmn accswap epi window
OUT=gpurun_out/${1:-r2}_mma_rate2.jsonl; : > $OUT
run() { timeout 30 ./tools/mma_rate "$@" >> $OUT 2>&1 || echo "{\"failed\": \"$*\"}" >> $OUT; }
run 4096 0 2 256 0 3 0 0 0 2     # baseline: SS, K-major B
run 4096 0 2 256 0 3 1 0 0 2     # MN-major B (row-major NN case)
run 4096 0 2 256 0 3 0 1 0 2     # accumulator swap every 4 k-blocks
run 4096 0 2 256 0 3 1 1 0 2
run 4096 0 2 256 0 3 0 1 1 2     # + continuous tcgen05.ld from 8 warps per CTA
run 4096 0 2 256 0 3 0 1 2 2     # + tcgen05.ld at ~1/8 duty
run 4096 0 2 256 0 3 1 1 2 2
run 4096 0 2 256 2 3 1 1 2 2     # + shared-memory noise
run 4096 0 2 256 2 3 1 1 2 3
run 4096 0 2 256 0 3 1 1 1 3
run 4096 1 2 256 0 3 1 1 2 2     # TS variants
run 4096 1 2 256 2 3 1 1 2 2
run 4096 0 1 128 0 3 0 0 0 2     # 1-CTA: window effect
run 4096 0 1 128 0 3 0 0 0 4
run 4096 0 1 128 0 3 0 0 0 6
run 4096 1 1 128 0 3 0 0 0 6
run 4096 0 2 128 0 3 0 0 0 6
run 4096 1 2 128 0 3 0 0 0 6
run 4096 1 2 128 2 3 1 0 0 6
run 4096 0 2 192 2 3 1 0 0 4
cat $OUT
