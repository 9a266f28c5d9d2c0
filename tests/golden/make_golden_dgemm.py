#!/usr/bin/env python
"""Generate tests/golden/ugemm_golden_dgemm.npz from the UNMODIFIED reference (dgemm_cpu ugemm.h:162, _dgemm_c
gemm_cpu.h:284 with real = double, dgemm_avx dgemm_avx.h:844).

    python tests/golden/make_golden_dgemm.py          # needs /root/reference compiled into oracle/_ref

Inputs are regenerated from the counter-based stream (tests/_oracle.py: make_problem_f64, seed 700 + case index); a CRC
of every input pins it.
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
    ("R", "N", "T", 129, 97, 131, 1.5, 0.5, (2, 2, 2), -0.5, 0.5),
    ("R", "T", "N", 129, 97, 131, 1.5, 0.5, (1, 2, 3), 0.0, 1.0),
    ("R", "T", "T", 129, 97, 131, -1.0, 2.0, (0, 0, 0), -0.5, 0.5),
    ("C", "N", "N", 65, 33, 47, 1.0, 1.0, (2, 0, 1), 0.0, 1.0),
    ("C", "T", "N", 65, 33, 47, 1.5, 0.5, (0, 3, 0), 0.0, 1.0),
    ("C", "N", "T", 65, 33, 47, 1.0, 0.0, (0, 0, 5), -0.5, 0.5),
    ("C", "T", "T", 65, 33, 47, 0.5, -1.0, (1, 1, 1), 0.0, 1.0),
    ("R", "N", "N", 1, 1, 1, 2.0, 3.0, (0, 0, 0), 0.0, 1.0),
    ("R", "N", "N", 256, 128, 64, 1.0, 0.0, (0, 0, 0), 0.0, 1.0),
]


def main():
    r = O.ref()
    if r is None or not hasattr(r, "ref_dgemm_cpu"):
        sys.exit("oracle/_ref/libugemm_ref.so missing or stale: run `make -C oracle ref` where /root/reference exists")
    out = {"cases": np.array([repr(c) for c in CASES])}
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(CASES):
        A, lda, B, ldb, Cm, ldc = O.make_problem_f64(maj, ta, tb, M, N, K, pad=pad, seed=700 + i, lo=lo, hi=hi)
        out[f"crc_{i}"] = np.array([zlib.crc32(A.tobytes()), zlib.crc32(B.tobytes()), zlib.crc32(Cm.tobytes())], dtype=np.uint64)
        fns = [("cpu", r.ref_dgemm_cpu), ("c", r.ref_dgemm_c)]
        if maj == "R":   # dgemm_avx segfaults on these column-major cases in the reference's own code [measured here]
            fns.append(("avx", r.ref_dgemm_avx))
        for name, fn in fns:
            out[f"{name}_{i}"] = O.run14(fn, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
    path = os.path.join(HERE, "ugemm_golden_dgemm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
