#!/bin/bash
# Bottleneck ablation of K1's 1-CTA tiles on tall-skinny shapes (UGEMM_K1_FLAGS, see csrc/common.cuh; results are WRONG under
# ablation flags, only the time matters):  1 = production; +2 transform skips its stores; +4 transform skips loads and stores;
# +8 only big*big is issued; +16 epilogue skips its stores; +64 MMA free-run (no TMA / transform / stage barriers).
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/skinny_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
for (M, N, K, ta) in ((200704, 64, 1152, "N"), (200704, 128, 1152, "N"), (200704, 64, 1152, "T"), (200704, 256, 1152, "N"), (4096, 4096, 4096, "N")):
    dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
    dA.fill_uniform(1); dB.fill_uniform(2)
    if M == 4096:
        u.set_k1_tuning(cta_group=1)
    lda = K if ta == "N" else M
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 10, 2, "R", ta, "N", M, N, K, 1.0, dA, lda, dB, N, 0.0, dC, N)
    out[f"{M}x{N}x{K}_{ta}N"] = round(best, 4)
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
PY
: > $OUT/skinny_ablation.jsonl
for F in 1 3 5 9 17 13 65; do
  UGEMM_K1_FLAGS=$F timeout 60 python /tmp/skinny_ab.py 2>/dev/null | tail -1 >> $OUT/skinny_ablation.jsonl
done
cat $OUT/skinny_ablation.jsonl
