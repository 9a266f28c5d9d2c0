#!/bin/bash
# Per-role cycle breakdown (UGEMM_K1_FLAGS bit 5) of K1 on the tall-skinny shape 200704 x 64 x 1152 and on c4 (N = 256).
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/skinny_once.py <<'PY'
import sys
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
M, K = 200704, 1152
for N in (64, 256):
    dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
    dA.fill_uniform(1); dB.fill_uniform(2)
    for _ in range(2):
        u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
        u.sync()
    print("N", N, u.sgemm_cuda_time_dev("3xtf32", 5, 1, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N), file=sys.stderr)
    dA.free(); dB.free(); dC.free()
PY
UGEMM_K1_FLAGS=$((1|32)) timeout 120 python /tmp/skinny_once.py 2> $OUT/skinny_prof.txt
grep -E "^N|cta0" $OUT/skinny_prof.txt | tail -8
