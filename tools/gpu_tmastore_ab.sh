#!/bin/bash
# A/B of K1's TMA-store epilogue (default) against the row-strided global stores (UGEMM_K1_FLAGS bit 13) on the BASELINE shapes.
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/ts_ab.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
for (M, N, K, ta, tb, beta, pad) in ((8192, 8192, 8192, "N", "N", 0.0, 0), (4095, 3001, 2047, "N", "T", 0.5, 1), (200704, 256, 1152, "N", "N", 0.0, 0), (200704, 256, 1152, "N", "N", 1.0, 0),
                                (4096, 4096, 4096, "N", "N", 0.0, 0), (2048, 2048, 2048, "N", "N", 0.0, 0), (200704, 128, 1152, "N", "N", 0.0, 0), (1024, 1024, 1024, "N", "N", 0.0, 0)):
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    lda, ldb, ldc = (ac + 3) // 4 * 4, (bc + 3) // 4 * 4, (N + 3) // 4 * 4
    dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", ta, tb, M, N, K, 1.0, dA, lda, dB, ldb, beta, dC, ldc)
    out[f"{M}x{N}x{K}_{ta}{tb}_b{beta}"] = [round(avg, 4), round(best, 4)]
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
PY
: > $OUT/tmastore_ab.jsonl
for F in 1 8193 1 8193; do
  UGEMM_K1_FLAGS=$F timeout 100 python /tmp/ts_ab.py 2>/dev/null | tail -1 >> $OUT/tmastore_ab.jsonl
done
cat $OUT/tmastore_ab.jsonl
