"""CPU suite: pins the oracle (oracle/sgemm_oracle.c) against the reference's known-answer vector, the
golden fixtures generated from the unmodified reference, and -- when oracle/_ref exists -- the reference
itself run live.  No GPU, no product code except the host RNG."""
import ast
import os
import zlib

import numpy as np
import pytest

import _oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ugemm_golden.npz")
TOL = 1e-6  # oracle vs reference implementations: same algorithm family, fp32, different summation order


def rel(x, r):
    return float(np.linalg.norm(x.astype(np.float64) - r.astype(np.float64)) / np.linalg.norm(r.astype(np.float64)))


def test_known_answer_vector():
    """sgemm_test.c:186-200: A=[[1,2],[3,4],[5,6]], B=[[1,2,3],[4,5,6]] -> [[9,12,15],[19,26,33],[29,40,51]]."""
    A = np.array([1, 2, 3, 4, 5, 6], np.float32)
    B = np.array([1, 2, 3, 4, 5, 6], np.float32)
    want = np.array([9, 12, 15, 19, 26, 33, 29, 40, 51], np.float32)
    C0 = np.zeros(9, np.float32)
    o = O.oracle()
    got = O.run14(o.oracle_sgemm_naive, "R", "N", "N", 3, 3, 2, 1.0, A, 2, B, 3, 0.0, C0, 3)
    assert np.array_equal(got, want)
    got = O.run14(o.oracle_sgemm_banded, "R", "N", "N", 3, 3, 2, 1.0, A, 2, B, 3, 0.0, C0, 3, threads=1)
    assert np.array_equal(got, want)
    r = O.ref()
    if r is not None:
        for fn in (r.ref_sgemm_cpu, r.ref_sgemm_c, r.ref_sgemm_avx, r.ref_sgemm_sse):
            assert np.array_equal(O.run14(fn, "R", "N", "N", 3, 3, 2, 1.0, A, 2, B, 3, 0.0, C0, 3), want)


def _golden_cases():
    g = np.load(GOLDEN)
    return g, [ast.literal_eval(str(c)) for c in g["cases"]]


def test_golden_inputs_are_reproducible():
    g, cases = _golden_cases()
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=100 + i, lo=lo, hi=hi)
        crc = [zlib.crc32(A.tobytes()), zlib.crc32(B.tobytes()), zlib.crc32(Cm.tobytes())]
        assert crc == [int(v) for v in g[f"crc_{i}"]], f"case {i}: RNG stream changed"


def test_oracle_matches_golden_reference_outputs():
    g, cases = _golden_cases()
    o = O.oracle()
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=100 + i, lo=lo, hi=hi)
        naive = O.run14(o.oracle_sgemm_naive, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
        band = O.run14(o.oracle_sgemm_banded, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=2)
        for name in ("cpu", "c", "sse", "avx"):
            key = f"{name}_{i}"
            if key not in g:
                continue
            want = g[key]
            assert rel(naive, want) <= TOL, (i, name, rel(naive, want))
            assert rel(band, want) <= TOL, (i, name, rel(band, want))
        # ld padding is never written by either oracle
        (ar, ac), (br, bc), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
        if pad[2]:
            assert np.array_equal(naive.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:])
            assert np.array_equal(band.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:])


def test_banded_oracle_is_bit_exact_with_golden_sgemm_avx_nn():
    """The 35-band restatement reproduces sgemm_avx's summation order exactly on tile-multiple NN cases."""
    g, cases = _golden_cases()
    o = O.oracle()
    hits = 0
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        if f"avx_{i}" not in g:
            continue
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=100 + i, lo=lo, hi=hi)
        band = O.run14(o.oracle_sgemm_banded, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=1)
        assert rel(band, g[f"avx_{i}"]) <= 2e-7
        hits += 1
    assert hits >= 3


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (reference tree absent)")
def test_oracle_matches_live_reference():
    o, r = O.oracle(), O.ref()
    for (maj, ta, tb) in [("R", "N", "N"), ("R", "N", "T"), ("R", "T", "N"), ("R", "T", "T"), ("C", "N", "N"), ("C", "T", "N")]:
        M, N, K = 255, 301, 207
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=(5, 3, 7), seed=7)
        want = O.run14(r.ref_sgemm_cpu, maj, ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        for fn, th in ((o.oracle_sgemm_naive, None), (o.oracle_sgemm_banded, 3)):
            got = O.run14(fn, maj, ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc, threads=th)
            assert rel(got, want) <= TOL
        sse = O.run14(r.ref_sgemm_sse, maj, ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        assert rel(sse, want) <= TOL
    # c1-like: 512^3 NN, banded == sgemm_avx bit for bit, threaded wrapper == single thread
    M = N = K = 512
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=1)
    avx = O.run14(r.ref_sgemm_avx, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    band = O.run14(o.oracle_sgemm_banded, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, threads=4)
    mt = O.run14(r.ref_sgemm_avx_mt, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, threads=4)
    assert np.array_equal(avx, band)
    assert np.array_equal(avx, mt)


def test_reference_quirks_are_restated():
    o = O.oracle()
    M, N, K = 5, 4, 3
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=3)
    # unknown trans letter: nothing written (ugemm.h: no matching branch)
    got = O.run14(o.oracle_sgemm_naive, "R", "X", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    assert np.array_equal(got, Cm)
    # naive propagates NaN through beta == 0 (0*NaN), banded overwrites (sgemm_avx256.h:324-330)
    Cn = Cm.copy()
    Cn[0] = np.nan
    assert np.isnan(O.run14(o.oracle_sgemm_naive, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cn, ldc)[0])
    assert not np.isnan(O.run14(o.oracle_sgemm_banded, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cn, ldc, threads=1)).any()
    # quick returns (sgemm_avx256.h:410): alpha == 0 and beta == 1 leaves C alone; alpha == 0 scales by beta
    assert np.array_equal(O.run14(o.oracle_sgemm_banded, "R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 1.0, Cm, ldc, threads=1), Cm)
    got = O.run14(o.oracle_sgemm_banded, "R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 0.5, Cm, ldc, threads=1)
    assert np.array_equal(got, (Cm * np.float32(0.5)).astype(np.float32))


def test_cmp_results_statistics():
    """oracle_cmp_results restates check_sgemm.c:56-85; compare with a direct numpy evaluation."""
    rng = np.random.default_rng(0)
    M, N, ld = 17, 13, 16
    ref = rng.uniform(0, 1, M * ld).astype(np.float32)
    res = ref.copy()
    res.reshape(M, ld)[3, 5] += np.float32(1e-3)
    out = np.zeros(4)
    verdict = O.oracle().oracle_cmp_results(M, N, ref, res, ld, out)
    r = ref.reshape(M, ld)[:, :N].astype(np.float64)
    x = res.reshape(M, ld)[:, :N].astype(np.float64)
    assert np.isclose(out[0], np.sqrt(((x - r) ** 2).mean()))
    assert np.isclose(out[1], r.std())
    assert np.isclose(out[2], np.abs(x - r).max())
    assert int(out[3]) == 3 * ld + 5
    assert verdict == 2  # maxErr > 1e-5 * stdRef => "FAIL !!!"
    assert O.oracle().oracle_cmp_results(M, N, ref, ref, ld, out) == 0


def test_relerr():
    ref = np.arange(1, 13, dtype=np.float32)
    res = ref.copy()
    res[2] += 1  # inside the 3x3 region of a ld=4 row-major matrix
    res[3] += 100  # padding: ignored
    e = O.relerr("R", 3, 3, ref, res, 4)
    want = 1.0 / np.sqrt(sum(float(v) ** 2 for v in (1, 2, 3, 5, 6, 7, 9, 10, 11)))
    assert np.isclose(e, want)


def test_host_rng_matches_oracle_bit_for_bit():
    import ugemm_b200 as u
    for seed, lo, hi in ((1, 0.0, 1.0), (2, -0.5, 0.5), (12345678901, 3.0, 7.0)):
        a = u.fill_uniform_host(4099, seed, lo, hi)
        b = O.fill_uniform(4099, seed, lo, hi)
        assert np.array_equal(a, b)
        assert a.min() >= lo and a.max() < hi + 1e-6
    x = O.fill_uniform(1 << 16, 5)
    assert abs(float(x.mean()) - 0.5) < 0.01
