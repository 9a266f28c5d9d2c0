#!/usr/bin/env python
"""DGEMM (K4) throughput on device-resident operands: 10 launches after 3 warm-ups, CUDA events.  JSON lines to gpurun_out/dgemm.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugemm_b200 as u  # noqa: E402

u.sgemm_cuda_init(0)
info = u.device_info()
peak = info["sm_count"] * 64 * 2 * info["sm_clock_khz"] * 1e3 / 1e12
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "dgemm.jsonl"), "a")
for ta, tb, M, N, K in (("N", "N", 4096, 4096, 4096), ("T", "N", 4096, 4096, 4096), ("N", "T", 4096, 4096, 4096), ("T", "T", 4096, 4096, 4096),
                        ("N", "N", 8192, 8192, 8192), ("N", "N", 1024, 1024, 1024), ("N", "N", 4095, 3001, 2047)):
    lda, ldb = (K if ta == "N" else M), (N if tb == "N" else K)
    ar, br = (M if ta == "N" else K), (K if tb == "N" else N)
    # fp64 buffers through the float allocator: 2 floats per double; contents = any finite bit patterns of moderate size
    dA, dB, dC = u.DeviceBuffer(2 * ar * lda), u.DeviceBuffer(2 * br * ldb), u.DeviceBuffer(2 * M * N)
    import numpy as np
    rng = np.random.default_rng(1)
    for d, n in ((dA, ar * lda), (dB, br * ldb)):
        h = rng.uniform(0, 1, n)
        u.lib().ugemm_cuda_memcpy_h2d(d.ptr, h.ctypes.data, h.nbytes)
    avg, best = u.dgemm_cuda_time_dev(10, 3, "R", ta, tb, M, N, K, 1.0, dA, lda, dB, ldb, 0.0, dC, N)
    rec = {"op": "dgemm", "ta": ta, "tb": tb, "M": M, "N": N, "K": K, "ms_avg": avg, "ms_min": best, "tflops_avg": 2.0 * M * N * K / avg / 1e9,
           "tflops_best": 2.0 * M * N * K / best / 1e9, "fp64_peak_tflops": peak, "frac": 2.0 * M * N * K / avg / 1e9 / peak}
    print(json.dumps(rec), flush=True)
    log.write(json.dumps(rec) + "\n")
    for d in (dA, dB, dC):
        d.free()
