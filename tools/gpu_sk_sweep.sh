#!/bin/bash
# TS kernel: cta_group x stream-K on/off over sizes that straddle the scheduling rules (one box, one process per flag set)
TAG=${1:-sk}
OUT=gpurun_out/${TAG}_sk_sweep.jsonl; : > $OUT
cat > /tmp/sk.py <<'PY'
import os, sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
flags = int(os.environ.get("UGEMM_K1_FLAGS", "1"))
shapes = [(512, 512, 4096), (1024, 1024, 1024), (1024, 1024, 8192), (1536, 1536, 1536), (2048, 2048, 2048), (2560, 2560, 2560), (3072, 3072, 3072), (4096, 3072, 2048),
          (4096, 4096, 4096), (5120, 5120, 5120), (6144, 6144, 6144), (8192, 8192, 1024), (200704, 256, 1152), (200704, 128, 1152), (200704, 64, 1152)]
for cg in (1, 2):
    u.set_k1_tuning(4, 0, cg)
    out = {"flags": flags, "cg": cg}
    for (M, N, K) in shapes:
        dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
        dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
        avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
        out[f"{M}x{N}x{K}"] = [round(avg * 1000, 1), round(2.0 * M * N * K / avg / 1e9, 1)]
        dA.free(); dB.free(); dC.free()
    print(json.dumps(out), flush=True)
PY
for F in $2; do UGEMM_K1_FLAGS=$F timeout 200 python /tmp/sk.py 2>&1 | grep flags >> $OUT; done
cat $OUT
