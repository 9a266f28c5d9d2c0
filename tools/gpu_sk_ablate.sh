#!/bin/bash
# what does the stream-K tail cost on 4096x3072x2048 / 4096^3?  TS kernel: SK on, SK on without fix-up pass, SK on without part stores, SK off
export UGEMM_K1_ABLATION=1
cat > /tmp/ska.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
for (M, N, K) in ((4096, 3072, 2048), (4096, 4096, 4096), (2560, 2560, 2560)):
    dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
    out[f"{M}x{N}x{K}"] = [round(avg * 1000, 1), round(best * 1000, 1)]
print(json.dumps(out))
PY
for F in $((1+131072)) $((1+131072+65536)) $((1+131072+65536+16)) 2049 $((2049+16)); do UGEMM_K1_FLAGS=$F timeout 100 python /tmp/ska.py 2>&1 | tail -1; done
