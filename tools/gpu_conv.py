#!/usr/bin/env python
"""Fused (implicit-GEMM) vs unfused (im2col + GEMM per image) convolution on the BASELINE config 4 layer: 64 images of
128 x 56 x 56, 256 filters 3 x 3, pad 1.  Wall clock around 10 launches + sync (each launch >= 0.5 ms).  JSON lines to gpurun_out/conv.jsonl."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugemm_b200 as u  # noqa: E402

u.sgemm_cuda_init(0)
log = open(os.path.join(ROOT, "gpurun_out", "conv.jsonl"), "a")
for nimg, ich, h, w, k, pad, ch in ((64, 128, 56, 56, 3, 1, 256), (16, 256, 28, 28, 3, 1, 512), (8, 64, 112, 112, 3, 1, 128), (1, 128, 56, 56, 3, 1, 256)):
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    dx = u.DeviceBuffer(nimg * ich * h * w).fill_uniform(1, -0.5, 0.5)
    dw = u.DeviceBuffer(ch * ich * k * k).fill_uniform(2, -0.5, 0.5)
    db = u.DeviceBuffer(ch).fill_uniform(3, -0.5, 0.5)
    dout = u.DeviceBuffer(nimg * ch * ho * wo)
    dws = u.DeviceBuffer(ich * k * k * ho * wo)
    flops = 2.0 * nimg * ch * ho * wo * ich * k * k
    for fusion in (1, 0):
        u.set_conv_fusion(fusion)
        for _ in range(2):
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, 1, dout, ch, db, 0.1, dws)
        u.sync()
        t0 = time.perf_counter()
        for _ in range(10):
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, 1, dout, ch, db, 0.1, dws)
        u.sync()
        ms = (time.perf_counter() - t0) * 100
        rec = {"op": "conv", "fused": bool(u.last_conv_fused()), "nimg": nimg, "ich": ich, "h": h, "w": w, "k": k, "ch": ch, "ms": ms, "tflops": flops / ms / 1e9,
               "col_matrix_bytes_avoided": 4 * nimg * ich * k * k * ho * wo if fusion else 0}
        print(json.dumps(rec), flush=True)
        log.write(json.dumps(rec) + "\n")
    u.set_conv_fusion(-1)
    for b in (dx, dw, db, dout, dws):
        b.free()
