#!/usr/bin/env python
"""Generate tests/golden/ugemm_golden.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled into oracle/_ref by oracle/Makefile):
    python tests/golden/make_golden.py
Inputs are NOT stored: they are regenerated from the counter-based stream
(oracle_fill_uniform, seed per case) -- a CRC of every input pins that stream.  Outputs stored per
case: sgemm_cpu (ugemm.h:287), sgemm_c (gemm_cpu.h:284), sgemm_sse (sgemm_sse.h:365) and, for
row-major NN only, sgemm_avx (sgemm_avx256.h:392; its T branches are empty).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

# (major, ta, tb, M, N, K, alpha, beta, (pad_a, pad_b, pad_c), lo, hi)
CASES = [
    ("R", "N", "N", 3, 3, 2, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
    ("R", "N", "N", 64, 96, 70, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
    ("R", "N", "N", 129, 97, 131, 1.5, 0.5, (3, 5, 7), 0.0, 1.0),
    ("R", "N", "T", 129, 97, 131, 1.5, 0.5, (4, 4, 4), -0.5, 0.5),
    ("R", "T", "N", 129, 97, 131, 1.5, 0.5, (1, 2, 3), 0.0, 1.0),
    ("R", "T", "T", 129, 97, 131, -1.0, 2.0, (0, 0, 0), -0.5, 0.5),
    ("C", "N", "N", 65, 33, 47, 1.0, 1.0, (2, 0, 1), 0.0, 1.0),
    ("C", "T", "N", 65, 33, 47, 1.5, 0.5, (0, 3, 0), 0.0, 1.0),
    ("C", "N", "T", 65, 33, 47, 1.0, 0.0, (0, 0, 5), -0.5, 0.5),
    ("C", "T", "T", 65, 33, 47, 0.5, -1.0, (1, 1, 1), 0.0, 1.0),
    ("R", "N", "N", 1, 1, 1, 2.0, 3.0, (0, 0, 0), 0.0, 1.0),
    ("R", "N", "N", 2, 200, 35, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
    ("R", "N", "N", 128, 64, 256, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
    ("R", "T", "N", 100, 72, 72, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
]


def main():
    r = O.ref()
    if r is None:
        sys.exit("oracle/_ref/libugemm_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    out = {"cases": np.array([repr(c) for c in CASES])}
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(CASES):
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=100 + i, lo=lo, hi=hi)
        out[f"crc_{i}"] = np.array([zlib.crc32(A.tobytes()), zlib.crc32(B.tobytes()), zlib.crc32(Cm.tobytes())],
                                   dtype=np.uint64)
        fns = {"cpu": r.ref_sgemm_cpu, "c": r.ref_sgemm_c, "sse": r.ref_sgemm_sse}
        if (maj, ta, tb) == ("R", "N", "N"):
            fns["avx"] = r.ref_sgemm_avx
        for name, fn in fns.items():
            out[f"{name}_{i}"] = O.run14(fn, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
    path = os.path.join(HERE, "ugemm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
