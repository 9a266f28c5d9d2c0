#!/usr/bin/env python
"""Generate tests/golden/ugemm_golden_l12.npz (SAXPY / SGEMV) from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled into oracle/_ref by oracle/Makefile):
    python tests/golden/make_golden_l12.py
Inputs are regenerated from the counter-based stream (oracle_fill_uniform, seed 200 + case index); a CRC of every
input pins that stream.  Outputs: saxpy_cpu (ugemm.h:75-86), sgemv_cpu (ugemm.h:124-150).  sgemv cases keep
incx == incy because the reference strides x by incy (ugemm.h:140,147).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

# (N, alpha, incx, incy)
AXPY = [(1, 2.0, 1, 1), (8, 0.5, 1, 1), (1000, -1.25, 1, 1), (1027, 0.5, 1, 1), (333, 3.0, 2, 3), (64, 1.0, 3, 1)]
# (trans, M, N, alpha, beta, lda_pad, inc)
GEMV = [("N", 1, 1, 2.0, 3.0, 0, 1), ("N", 65, 33, 1.0, 0.0, 0, 1), ("T", 65, 33, 1.5, 0.5, 0, 1), ("N", 129, 257, 1.5, 0.5, 3, 1),
        ("T", 129, 257, -1.0, 2.0, 4, 1), ("N", 40, 70, 1.0, 1.0, 1, 2), ("T", 40, 70, 0.5, -1.0, 2, 3), ("T", 7, 1024, 1.0, 0.0, 0, 1)]


def axpy_inputs(i, N, incx, incy):
    return O.fill_uniform((N - 1) * incx + 1, 200 + 2 * i, -1.0, 1.0), O.fill_uniform((N - 1) * incy + 1, 201 + 2 * i, -1.0, 1.0)


def gemv_inputs(i, trans, M, N, lda_pad, inc):
    lines, cols = (N, M) if trans == "N" else (M, N)
    lda = cols + lda_pad
    A = O.fill_uniform(lines * lda, 300 + 3 * i, 0.0, 1.0)
    x = O.fill_uniform((N - 1) * inc + 1, 301 + 3 * i, -0.5, 0.5)
    y = O.fill_uniform((M - 1) * inc + 1, 302 + 3 * i, 0.0, 1.0)
    return A, lda, x, y


def main():
    r = O.ref()
    if r is None or not hasattr(r, "ref_saxpy_cpu"):
        sys.exit("oracle/_ref/libugemm_ref.so missing or stale: run `make -C oracle ref` where /root/reference exists")
    out = {"axpy": np.array([repr(c) for c in AXPY]), "gemv": np.array([repr(c) for c in GEMV])}
    for i, (N, alpha, incx, incy) in enumerate(AXPY):
        x, y = axpy_inputs(i, N, incx, incy)
        out[f"axpy_crc_{i}"] = np.array([zlib.crc32(x.tobytes()), zlib.crc32(y.tobytes())], dtype=np.uint64)
        r.ref_saxpy_cpu(N, alpha, x, incx, y, incy)
        out[f"axpy_{i}"] = y
    for i, (trans, M, N, alpha, beta, lda_pad, inc) in enumerate(GEMV):
        A, lda, x, y = gemv_inputs(i, trans, M, N, lda_pad, inc)
        out[f"gemv_crc_{i}"] = np.array([zlib.crc32(A.tobytes()), zlib.crc32(x.tobytes()), zlib.crc32(y.tobytes())], dtype=np.uint64)
        r.ref_sgemv_cpu(trans.encode(), M, N, alpha, A, lda, x, inc, beta, y, inc)
        out[f"gemv_{i}"] = y
    path = os.path.join(HERE, "ugemm_golden_l12.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
