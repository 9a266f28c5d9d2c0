#!/usr/bin/env python
"""Generate tests/golden/ugemm_golden_conv.npz (im2col / convolution / convolution + bias + LeakyReLU) from the
UNMODIFIED reference.

Run in the build container (needs /root/reference; oracle/Makefile compiles sgemm_gl1.h where it lies against the GL/GLFW
stand-ins of oracle/stubs/ into oracle/_ref/libugemm_ref_conv.so):
    python tests/golden/make_golden_conv.py
Inputs are regenerated from the counter-based stream (oracle_fill_uniform, seeds 500 + 3*case ...); a CRC of every input
pins that stream.  Outputs per case: `col` (+ its CRC; the CRC alone for the larger cases) = the reference's CPU im2col (sgemm_gl1.h:166-190), `out_cpu` / `out_sse` = that
column matrix multiplied by the reference's sgemm_cpu / sgemm_sse in gl_convolution_LReLU's orientation (M = ch,
N = hcol*wcol, K = k*k*ich, sgemm_gl1.h:200-207), `act_sse` = the same plus the bias + LeakyReLU(0.1) loop of
sgemm_gl1.h:210-217.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

# ich, h, w, k, pad, stride, ch
CASES = [
    (1, 5, 5, 3, 1, 1, 2),
    (3, 8, 10, 3, 0, 1, 4),
    (4, 9, 7, 2, 1, 2, 5),
    (8, 14, 14, 3, 1, 1, 16),
    (2, 6, 6, 5, 2, 1, 3),
    (3, 12, 9, 4, 2, 2, 6),
    (6, 7, 9, 3, 2, 1, 8),          # pad > (k-1)/2: output larger than the input
    (32, 8, 32, 3, 1, 1, 64),       # smallest geometry the fused implicit-GEMM path takes
    (16, 16, 16, 1, 0, 1, 24),      # 1x1
    (24, 12, 28, 5, 2, 3, 40),      # 5x5, stride 3
]


def inputs(i, ich, h, w, k, ch):
    return (O.fill_uniform(ich * h * w, 500 + 3 * i, -0.5, 0.5), O.fill_uniform(ch * ich * k * k, 501 + 3 * i, -0.5, 0.5),
            O.fill_uniform(ch, 502 + 3 * i, -0.5, 0.5))


def main():
    r = O.ref_conv()
    if r is None:
        sys.exit("oracle/_ref/libugemm_ref_conv.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {"cases": np.array([repr(c) for c in CASES])}
    for i, (ich, h, w, k, pad, stride, ch) in enumerate(CASES):
        x, wgt, bias = inputs(i, ich, h, w, k, ch)
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        out[f"crc_{i}"] = np.array([zlib.crc32(x.tobytes()), zlib.crc32(wgt.tobytes()), zlib.crc32(bias.tobytes())], dtype=np.uint64)
        col = np.full(ich * k * k * ho * wo, np.nan, np.float32)
        r.ref_im2col(x, ich, h, w, k, k, pad, pad, stride, stride, col)
        out[f"col_crc_{i}"] = np.array([zlib.crc32(col.tobytes())], dtype=np.uint64)   # im2col is data movement: a CRC pins it bit for bit
        if col.size <= 20000:
            out[f"col_{i}"] = col
        for name, which, bp in (("out_cpu", 0, None), ("out_sse", 1, None), ("act_sse", 1, bias.ctypes.data)):
            if name == "out_cpu" and ch * ho * wo > 4000:
                continue                      # the naive-GEMM twin only for the small cases (keeps the fixture small)
            o = np.full(ch * ho * wo, np.nan, np.float32)
            assert r.ref_gl_convolution(which, x, ich, w, h, wgt, k, pad, stride, o, ch, bp) == 0
            out[f"{name}_{i}"] = o
    path = os.path.join(HERE, "ugemm_golden_conv.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
