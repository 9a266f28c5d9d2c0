#!/usr/bin/env python
"""A/B of K1's narrow-N instantiation on tall-skinny shapes (GPU box):  K1 narrow (default) vs K1 with the full
128-column tile (UGEMM_K1_FLAGS bit 12) vs K2, with a sampled fp64 check of every K1 result.

    python tools/gpu_narrow.py            # spawns one subprocess per flag setting, appends to gpurun_out/narrow.jsonl
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = os.path.join(OUT, "narrow.jsonl")

SHAPES = [(200704, n, 1152) for n in (8, 16, 32, 48, 64, 96, 112, 128)] + [(8192, 64, 8192), (65536, 48, 512), (4096, 96, 4096)]


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(LOG, "a") as f:
        f.write(line + "\n")


def worker(tag):
    import ugemm_b200 as u
    u.sgemm_cuda_init()
    for (M, N, K) in SHAPES:
        for tb in ("N", "T"):
            ldb = N if tb == "N" else K
            dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
            dA.fill_uniform(1, -0.5, 0.5)
            dB.fill_uniform(2, -0.5, 0.5)
            dC.fill_uniform(3, 0.0, 1.0)
            for mode in (("3xtf32",) if tag != "narrow" else ("3xtf32", "simt")):
                if mode == "3xtf32" and (ldb % 4 or K % 4):
                    continue
                try:
                    u.sgemm_cuda_dev(mode, None, "R", "N", tb, M, N, K, 1.0, dA, K, dB, ldb, 0.0, dC, N)
                    u.sync()
                    rows = np.r_[0:512, M - 512:M] if M > 1024 else np.arange(M)
                    A = np.concatenate([dA.download(512 * K, 0), dA.download(512 * K, (M - 512) * K)]).reshape(-1, K) if M > 1024 else dA.download().reshape(M, K)
                    B = dB.download().reshape((K, N) if tb == "N" else (N, K))
                    opB = B if tb == "N" else B.T
                    got = np.concatenate([dC.download(512 * N, 0), dC.download(512 * N, (M - 512) * N)]).reshape(-1, N) if M > 1024 else dC.download().reshape(M, N)
                    ref = A.astype(np.float64) @ opB.astype(np.float64)
                    rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
                    avg, best = u.sgemm_cuda_time_dev(mode, 20, 3, "R", "N", tb, M, N, K, 1.0, dA, K, dB, ldb, 0.0, dC, N)
                    emit(tag=tag, mode=mode, kernel=u.last_kernel(), M=M, N=N, K=K, tb=tb, ms_avg=avg, ms_min=best, relerr=rel,
                         a_gbs=M * K * 4 / best / 1e6, tflops=2.0 * M * N * K / best / 1e9, rows_checked=int(rows.size))
                except Exception as e:  # noqa: BLE001
                    emit(tag=tag, mode=mode, M=M, N=N, K=K, tb=tb, error=str(e)[:300])
                    return 1
            dA.free(); dB.free(); dC.free()
    return 0


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(worker(sys.argv[1]))
    for tag, flags in (("narrow", "1"), ("full_tile", str(1 | 4096))):
        env = dict(os.environ, UGEMM_K1_FLAGS=flags)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=env, timeout=240)
            emit(tag=tag, rc=r.returncode)
        except subprocess.TimeoutExpired:
            emit(tag=tag, rc="timeout")
