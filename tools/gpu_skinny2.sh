#!/bin/bash
# 200704 x N x 1152 for small N: K1 (forced 3xtf32, single-CTA tiles) against K2 (forced simt)
cat > /tmp/sk2.py <<'PY'
import sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {}
M, K = 200704, 1152
dA = u.DeviceBuffer(M * K).fill_uniform(1)
for N in (8, 16, 32, 48, 64, 96, 128, 192, 256):
    dB, dC = u.DeviceBuffer(K * N).fill_uniform(2), u.DeviceBuffer(M * N)
    row = {}
    for mode in ("3xtf32", "simt"):
        try:
            avg, best = u.sgemm_cuda_time_dev(mode, 10, 3, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
            row[mode] = round(avg, 4)
        except Exception as ex:
            row[mode] = None
            u.backend.lib().sgemm_cuda_clear_error()
    out[N] = row
    dB.free(); dC.free()
print(json.dumps(out))
PY
timeout 200 python /tmp/sk2.py 2>&1 | tail -1
