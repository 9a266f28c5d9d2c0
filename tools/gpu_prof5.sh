#!/bin/bash
# per-role cycles (PROF build) of the TS kernel on 4096x3072x2048 with and without beta, stream-K off
TAG=${1:-r2k}
export UGEMM_K1_ABLATION=1
cat > /tmp/p5.py <<'PY'
import os, sys
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
M, N, K = 4096, 3072, 2048
dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(N * K), u.DeviceBuffer(M * N)
dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
beta = float(sys.argv[1])
print("beta", beta, flush=True)
avg, best = u.sgemm_cuda_time_dev("3xtf32", 2, 1, "R", "N", "T", M, N, K, 1.5, dA, K, dB, K, beta, dC, N)
PY
for b in 0.0 0.5; do UGEMM_K1_FLAGS=$((2048+32+1)) timeout 60 python /tmp/p5.py $b 2>&1 | grep -E "beta|k1prof cta[01] " | tail -3 | cut -c1-400; done
