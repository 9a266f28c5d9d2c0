#!/usr/bin/env python
"""Fused (implicit-GEMM) convolution only, three layers; CUDA-event-free wall clock around 20 launches + sync.  JSON lines to stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ugemm_b200 as u
u.sgemm_cuda_init(0)
u.set_conv_fusion(1)
for nimg, ich, h, w, k, pad, ch in ((64, 128, 56, 56, 3, 1, 256), (16, 256, 28, 28, 3, 1, 512), (8, 64, 112, 112, 3, 1, 128)):
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    dx = u.DeviceBuffer(nimg * ich * h * w).fill_uniform(1, -0.5, 0.5)
    dw = u.DeviceBuffer(ch * ich * k * k).fill_uniform(2, -0.5, 0.5)
    db = u.DeviceBuffer(ch).fill_uniform(3, -0.5, 0.5)
    dout = u.DeviceBuffer(nimg * ch * ho * wo)
    flops = 2.0 * nimg * ch * ho * wo * ich * k * k
    best = 1e9
    for rep in range(3):
        for _ in range(2):
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, 1, dout, ch, db, 0.1, None)
        u.sync(); time.sleep(0.3)
        t0 = time.perf_counter()
        for _ in range(20):
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, 1, dout, ch, db, 0.1, None)
        u.sync()
        best = min(best, (time.perf_counter() - t0) * 50)
    print(json.dumps({"op": "conv fused", "fused": bool(u.last_conv_fused()), "nimg": nimg, "ich": ich, "h": h, "w": w, "ch": ch, "ms": round(best, 4), "tflops": round(flops / best / 1e9, 1)}), flush=True)
    for b in (dx, dw, db, dout):
        b.free()
