"""DGEMM (the path of check_dgemm.c).

CPU part: oracle_dgemm_naive against the golden outputs of the unmodified reference (dgemm_cpu / _dgemm_c / dgemm_avx,
tests/golden/make_golden_dgemm.py) and against the live _ref.
GPU part (-m gpu): dgemm_cuda through the C ABI against the oracle on the same seeded inputs.  Gate: normwise relative
error <= 2e-14 -- the fp64 analogue of the SGEMM gate (1e-5 is 168 units of fp32 round-off; 168 * 2^-53 = 1.9e-14) --
with inputs whose low mantissa bits are populated, so a kernel that computed in fp32 anywhere would miss it by 1e6.
"""
import ast
import os
import zlib

import numpy as np
import pytest

import _oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ugemm_golden_dgemm.npz")
TOL64 = 2e-14


def rel64(lines, cols, ref, res, ld):
    return float(O.oracle().oracle_relerr_f64(lines, cols, ref, res, ld))


def _cases():
    g = np.load(GOLDEN)
    return g, [ast.literal_eval(str(c)) for c in g["cases"]]


def test_oracle_dgemm_matches_golden_reference_outputs():
    g, cases = _cases()
    o = O.oracle()
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        A, lda, B, ldb, Cm, ldc = O.make_problem_f64(maj, ta, tb, M, N, K, pad=pad, seed=700 + i, lo=lo, hi=hi)
        assert [zlib.crc32(A.tobytes()), zlib.crc32(B.tobytes()), zlib.crc32(Cm.tobytes())] == [int(v) for v in g[f"crc_{i}"]]
        mine = O.run14(o.oracle_dgemm_naive, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
        (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
        for name, tol in (("cpu", 1e-16), ("c", 1e-15), ("avx", 1e-15)):   # same order (one FMA-contraction choice apart) / blocked orders
            if f"{name}_{i}" in g:
                assert rel64(cr, cc, g[f"{name}_{i}"], mine, ldc) <= tol, (i, name)
        if pad[2]:
            assert np.array_equal(mine.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:])


def test_oracle_dgemm_matches_live_reference():
    r = O.ref()
    if r is None or not hasattr(r, "ref_dgemm_cpu"):
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    o = O.oracle()
    for ta in "NT":
        for tb in "NT":
            A, lda, B, ldb, Cm, ldc = O.make_problem_f64("R", ta, tb, 70, 50, 90, pad=(1, 2, 3), seed=9)
            mine = O.run14(o.oracle_dgemm_naive, "R", ta, tb, 70, 50, 90, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
            for fn in (r.ref_dgemm_cpu, r.ref_dgemm_c, r.ref_dgemm_avx):
                assert rel64(70, 50, O.run14(fn, "R", ta, tb, 70, 50, 90, 1.5, A, lda, B, ldb, 0.5, Cm, ldc), mine, ldc) <= 1e-15


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def u():
    import ugemm_b200 as u
    u.sgemm_cuda_init(0)
    return u


def gpu_case(u, maj, ta, tb, M, N, K, alpha, beta, pad, seed, lo=0.0, hi=1.0):
    A, lda, B, ldb, Cm, ldc = O.make_problem_f64(maj, ta, tb, M, N, K, pad=pad, seed=seed, lo=lo, hi=hi, sentinel=-77.0)
    got = Cm.copy()
    u.dgemm_cuda(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, got, ldc)
    want = O.run14(O.oracle().oracle_dgemm_naive, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
    (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
    e = rel64(cr, cc, want, got, ldc)
    assert e <= TOL64, f"dgemm {maj}{ta}{tb} {M}x{N}x{K} a={alpha} b={beta} pad={pad}: relerr {e:.3e} > {TOL64}"
    if pad[2] and cr * ldc:
        assert np.array_equal(got.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:]), "ld padding of C was written"
    return e


@pytest.mark.gpu
def test_dgemm_golden_fixtures(u):
    g, cases = _cases()
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        A, lda, B, ldb, Cm, ldc = O.make_problem_f64(maj, ta, tb, M, N, K, pad=pad, seed=700 + i, lo=lo, hi=hi)
        got = Cm.copy()
        u.dgemm_cuda(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, got, ldc)
        (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
        for name in ("cpu", "c", "avx"):
            if f"{name}_{i}" in g:
                assert rel64(cr, cc, g[f"{name}_{i}"], got, ldc) <= TOL64, (i, name)
        if pad[2]:
            assert np.array_equal(got.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:])


@pytest.mark.gpu
@pytest.mark.parametrize("maj", ["R", "C"])
@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_dgemm_sweep_vs_oracle(u, maj, ta, tb):
    """tiny, ragged, tile-multiple and odd-ld shapes: edge tiles, the interior fast path, scalar and 128-bit layouts"""
    worst = 0.0
    for i, (M, N, K) in enumerate(((1, 1, 1), (2, 3, 5), (17, 9, 33), (128, 64, 8), (128, 64, 9), (129, 65, 16), (256, 128, 64), (300, 200, 100),
                                   (64, 300, 257), (513, 130, 70))):
        for alpha, beta, pad in ((1.0, 0.0, (0, 0, 0)), (1.5, 0.5, (2, 4, 6)), (-1.0, 2.0, (1, 3, 5))):
            worst = max(worst, gpu_case(u, maj, ta, tb, M, N, K, alpha, beta, pad, seed=20 + i))
    print(f"dgemm {maj}{ta}{tb}: worst relerr {worst:.3e}")


@pytest.mark.gpu
def test_dgemm_semantics_and_errors(u):
    M, N, K = 40, 30, 20
    A, lda, B, ldb, Cm, ldc = O.make_problem_f64("R", "N", "N", M, N, K, seed=3)
    got = np.full(M * N, np.nan)
    u.dgemm_cuda("R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, got, ldc)      # beta == 0 never reads C
    assert np.isfinite(got).all()
    got = Cm.copy(); u.dgemm_cuda("R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 0.5, got, ldc)
    assert np.array_equal(got, 0.5 * Cm)
    got = Cm.copy(); u.dgemm_cuda("R", "N", "N", M, N, 0, 1.0, A, 1, B, ldb, 1.0, got, ldc)
    assert np.array_equal(got, Cm)
    got = Cm.copy(); u.dgemm_cuda("r", "n", "n", M, N, K, 1.0, A, lda, B, ldb, 0.0, got, ldc)   # lower case accepted
    assert rel64(M, N, A.reshape(M, K) @ B.reshape(K, N), got, ldc) <= TOL64
    for bad in (lambda c: u.dgemm_cuda("X", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, c, ldc),
                lambda c: u.dgemm_cuda("R", "Q", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, c, ldc),
                lambda c: u.dgemm_cuda("R", "N", "N", M, N, K, 1.0, A, K - 1, B, ldb, 0.0, c, ldc)):
        got = Cm.copy()
        with pytest.raises(u.UgemmCudaError):
            bad(got)
        assert np.array_equal(got, Cm)
    with pytest.raises(TypeError):   # float32 buffers are not silently reinterpreted
        u.dgemm_cuda("R", "N", "N", M, N, K, 1.0, A.astype(np.float32), lda, B, ldb, 0.0, Cm.copy(), ldc)


@pytest.mark.gpu
def test_dgemm_check_dgemm_shape_and_linearity(u):
    """check_dgemm.c's default-sized case vs the oracle (1024^3 would take the naive oracle ~10 s; 512 x 512 x 1024 here),
    then 4096^3 on the device through a size-independent property: A(B1 + B2) = A B1 + A B2 to fp64 round-off."""
    gpu_case(u, "R", "N", "N", 512, 512, 1024, 1.0, 0.0, (0, 0, 0), seed=77)
    n = 4096
    x = O.fill_uniform(n * n, 91).astype(np.float64) + O.fill_uniform(n * n, 92).astype(np.float64) * 2.0 ** -24
    b1 = O.fill_uniform(n * n, 93, -0.5, 0.5).astype(np.float64)
    b2 = O.fill_uniform(n * n, 94, -0.5, 0.5).astype(np.float64) * (1 + 2.0 ** -30)

    class DBuf:
        def __init__(self, host):
            self.b = u.DeviceBuffer(2 * host.size)
            u.lib().ugemm_cuda_memcpy_h2d(self.b.ptr, host.ctypes.data, host.nbytes)

        def data_ptr(self):
            return self.b.ptr

        def download(self):
            out = np.empty(self.b.n // 2, np.float64)
            u.lib().ugemm_cuda_memcpy_d2h(out.ctypes.data, self.b.ptr, out.nbytes)
            return out

    dA, d1, d2, ds = DBuf(x), DBuf(b1), DBuf(b2), DBuf(b1 + b2)
    c1, c2, cs = DBuf(np.zeros(n * n)), DBuf(np.zeros(n * n)), DBuf(np.zeros(n * n))
    for db, dc in ((d1, c1), (d2, c2), (ds, cs)):
        u.dgemm_cuda_dev(None, "R", "N", "N", n, n, n, 1.0, dA, n, db, n, 0.0, dc, n)
    u.sync()
    s, a, b = cs.download(), c1.download(), c2.download()
    scale = np.sqrt(n) * 0.25 * n   # typical |term| magnitude: sum of n products of size <= 0.5, i.e. fp64 round-off reference
    assert np.abs(s - (a + b)).max() <= 1e-12 * scale
    # one sampled row against fp64 numpy
    r = 1234
    want = x.reshape(n, n)[r] @ b1.reshape(n, n)
    assert np.linalg.norm(a.reshape(n, n)[r] - want) / np.linalg.norm(want) <= TOL64
    avg, best = u.dgemm_cuda_time_dev(3, 1, "R", "N", "N", n, n, n, 1.0, dA, n, d1, n, 0.0, c1, n)
    print(f"dgemm 4096^3: {2.0 * n ** 3 / best / 1e9:.1f} TFLOP/s best, {2.0 * n ** 3 / avg / 1e9:.1f} avg")
