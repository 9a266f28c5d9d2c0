#!/bin/bash
# Generic A/B of K1 flag sets on the BASELINE shapes (one process per flag set, interleaved twice so box drift shows).
# usage under gpurun: bash tools/gpu_ab.sh <tag> "<flags1> <flags2> ..." [extra env assignments]
TAG=${1:-ab}; FLAGS=${2:-"1 16385"}
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/ab.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
shapes = ((8192, 8192, 8192, "N", "N", 0.0), (4095, 3001, 2047, "N", "T", 0.5), (4095, 3001, 2047, "T", "N", 0.5), (200704, 256, 1152, "N", "N", 0.0),
          (4096, 4096, 4096, "N", "N", 0.0), (2048, 2048, 2048, "N", "N", 0.0), (200704, 128, 1152, "N", "N", 0.0), (1024, 1024, 1024, "N", "N", 0.0))
for (M, N, K, ta, tb, beta) in shapes:
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    lda, ldb, ldc = (ac + 3) // 4 * 4, (bc + 3) // 4 * 4, (N + 3) // 4 * 4
    dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", ta, tb, M, N, K, 1.0, dA, lda, dB, ldb, beta, dC, ldc)
    out[f"{M}x{N}x{K}_{ta}{tb}"] = [round(avg, 4), round(best, 4), round(2.0 * M * N * K / avg / 1e9, 1)]
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
PY
: > $OUT/${TAG}_ab.jsonl
for rep in 1 2; do for F in $FLAGS; do
  env $3 UGEMM_K1_FLAGS=$F timeout 100 python /tmp/ab.py 2>&1 | tail -1 >> $OUT/${TAG}_ab.jsonl
done; done
cat $OUT/${TAG}_ab.jsonl
