#!/usr/bin/env python
"""Achieved HBM GB/s of the level-1/2 kernels (device-resident operands larger than L2, 20 back-to-back launches
bracketed by a device sync; launch overhead ~5 us against 0.2-1 ms kernels).  Appends JSON lines to gpurun_out/l12.jsonl."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugemm_b200 as u  # noqa: E402

PEAK = 6448.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = os.path.join(ROOT, "gpurun_out", "l12.jsonl")


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    u.sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    u.sync()
    return (time.perf_counter() - t0) / iters


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    open(LOG, "a").write(line + "\n")


u.sgemm_cuda_init(0)
for n in (1 << 24, 1 << 28):
    dx, dy = u.DeviceBuffer(n).fill_uniform(1), u.DeviceBuffer(n).fill_uniform(2)
    s = timed(lambda: u.saxpy_cuda_dev(None, n, 0.5, dx, 1, dy, 1))
    emit(op="saxpy", n=n, ms=s * 1e3, gbs=12 * n / s / 1e9, frac_of_hbm_peak=12 * n / s / 1e9 / PEAK)
    dx.free(); dy.free()
for trans, M, N in (("T", 16384, 16384), ("N", 16384, 16384), ("T", 262144, 1024), ("N", 262144, 1024), ("T", 1024, 262144), ("N", 1024, 262144),
                    ("T", 200704, 1152), ("N", 4096, 4096), ("T", 4096, 4096)):
    lda = M if trans == "N" else N
    dA = u.DeviceBuffer(M * N).fill_uniform(3)
    dx, dy = u.DeviceBuffer(N).fill_uniform(4, -0.5, 0.5), u.DeviceBuffer(M)
    s = timed(lambda: u.sgemv_cuda_dev(None, trans, M, N, 1.0, dA, lda, dx, 1, 0.0, dy, 1))
    b = 4.0 * (M * N + M + N)
    emit(op="sgemv", trans=trans, M=M, N=N, ms=s * 1e3, gbs=b / s / 1e9, frac_of_hbm_peak=b / s / 1e9 / PEAK, gflops=2.0 * M * N / s / 1e9)
    dA.free(); dx.free(); dy.free()
