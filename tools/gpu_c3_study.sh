#!/bin/bash
# where does config 3 (4095x3001x2047, alpha=1.5 beta=0.5) lose its time?  one factor at a time; usage: bash tools/gpu_c3_study.sh <tag> <base flags>
TAG=${1:-c3}; BASE=${2:-1}
OUT=gpurun_out/${TAG}_c3_study.jsonl; : > $OUT
cat > /tmp/c3.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
cases = [("c3 NT b.5", 4095, 3001, 2047, "N", "T", 1.5, 0.5), ("c3 NT b0", 4095, 3001, 2047, "N", "T", 1.5, 0.0), ("c3 NN b0", 4095, 3001, 2047, "N", "N", 1.5, 0.0),
         ("c3 TN b0", 4095, 3001, 2047, "T", "N", 1.5, 0.0), ("K2048 NT b0", 4095, 3001, 2048, "N", "T", 1.0, 0.0), ("4096x3072x2048 NT b0", 4096, 3072, 2048, "N", "T", 1.0, 0.0),
         ("4096x3072x2048 NN b0", 4096, 3072, 2048, "N", "N", 1.0, 0.0), ("4736x4096x2048 NN (148 tiles)", 4736, 4096, 2048, "N", "N", 1.0, 0.0),
         ("4736x4096x4096 NN (148 tiles)", 4736, 4096, 4096, "N", "N", 1.0, 0.0), ("4736x2048x2048 NN (74 tiles)", 4736, 2048, 2048, "N", "N", 1.0, 0.0)]
for (name, M, N, K, ta, tb, alpha, beta) in cases:
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    lda, ldb, ldc = (ac + 3) // 4 * 4, (bc + 3) // 4 * 4, (N + 3) // 4 * 4
    dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc)
    out[name] = [round(avg, 4), round(best, 4), round(2.0 * M * N * K / avg / 1e9, 1)]
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
PY
for F in $BASE $((BASE + 2048)) $3; do
  UGEMM_K1_FLAGS=$F timeout 100 python /tmp/c3.py 2>&1 | tail -1 >> $OUT
done
cat $OUT
