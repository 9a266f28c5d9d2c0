"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the oracle on identical seeded
inputs.  Gate (north_star): normwise relative error ||C - C_ref||_F / ||C_ref||_F <= 1e-5 for both kernels;
the a-priori bound K*u (u = 2^-24) and the probabilistic sqrt(K)*u are printed per shape.  ld padding must be
bit-identical before/after.  Full-size BASELINE shapes are checked on sampled row slabs regenerated from the
shared counter-based RNG (the CPU cannot recompute 8192^3 in seconds)."""
import ast
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5
U = 2.0 ** -24
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ugemm_golden.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bounds(K):
    return f"K*u={K * U:.2e} sqrt(K)*u={np.sqrt(K) * U:.2e}"


def fn_of(u, mode):
    return {"auto": u.sgemm_cuda, "3xtf32": u.sgemm_cuda_3xtf32, "simt": u.sgemm_cuda_simt}[mode]


def gpu14(u, mode, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    out = Cm.copy()
    fn_of(u, mode)(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, out, ldc)
    return out


def oracle14(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, naive=False):
    o = O.oracle()
    if naive:
        return O.run14(o.oracle_sgemm_naive, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
    return O.run14(o.oracle_sgemm_banded, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc,
                   threads=min(8, O.oracle().oracle_max_threads()))


def check_case(u, mode, maj, ta, tb, M, N, K, alpha, beta, pad, seed=1, lo=0.0, hi=1.0, naive=False, tol=TOL):
    A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=seed, lo=lo, hi=hi, sentinel=-77.0)
    got = gpu14(u, mode, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
    want = oracle14(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, naive=naive)
    (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
    e = O.relerr("R", cr, cc, want, got, ldc)  # stored layout: cr lines of cc elements, pitch ldc
    assert e <= tol, f"{mode} {maj}{ta}{tb} {M}x{N}x{K} a={alpha} b={beta} pad={pad}: relerr {e:.3e} > {tol} ({bounds(K)})"
    if pad[2] and cr * ldc:
        assert np.array_equal(got.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:]), "ld padding of C was written"
    return e


def test_known_answer_vector(u):
    """sgemm_test.c:186-200"""
    A = np.array([1, 2, 3, 4, 5, 6], np.float32)
    B = np.array([1, 2, 3, 4, 5, 6], np.float32)
    want = np.array([9, 12, 15, 19, 26, 33, 29, 40, 51], np.float32)
    for mode in ("auto", "simt"):
        Cm = np.zeros(9, np.float32)
        fn_of(u, mode)("R", "N", "N", 3, 3, 2, 1.0, A, 2, B, 3, 0.0, Cm, 3)
        assert np.array_equal(Cm, want)
    assert u.last_kernel() == "simt"
    Cm = np.zeros(9, np.float32)
    u.sgemm_rnn(3, 3, 2, 1.0, A, B, 0.0, Cm)   # macro API of sgemm_test.c:19-33
    assert np.array_equal(Cm, want)


def test_device_rng_matches_host_and_oracle(u):
    for seed, lo, hi in ((1, 0.0, 1.0), (9, -0.5, 0.5)):
        n = 1 << 20
        d = u.DeviceBuffer(n).fill_uniform(seed, lo, hi)
        got = d.download()
        assert np.array_equal(got, O.fill_uniform(n, seed, lo, hi))
        assert np.array_equal(got, u.fill_uniform_host(n, seed, lo, hi))
        d.free()


SMALL_DIMS = [1, 2, 3, 7, 8, 31, 33, 64, 95, 97, 129]


@pytest.mark.parametrize("maj", ["R", "C"])
@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_k2_small_sweep_vs_naive_oracle(u, maj, ta, tb):
    """K2 on tiny/odd shapes, every transpose and major, tight and odd-padded ld, against the restated
    sgemm_cpu (ugemm.h:287)."""
    rng = np.random.default_rng(hash((maj, ta, tb)) % 2 ** 31)
    combos = [(1.0, 0.0), (1.5, 0.5), (1.0, 1.0), (-1.0, 2.0)]
    worst = 0.0
    for i in range(40):
        M, N, K = (int(rng.choice(SMALL_DIMS)) for _ in range(3))
        alpha, beta = combos[i % len(combos)]
        pad = (0, 0, 0) if i % 2 == 0 else (int(rng.integers(1, 6)), int(rng.integers(1, 6)), int(rng.integers(1, 6)))
        worst = max(worst, check_case(u, "simt", maj, ta, tb, M, N, K, alpha, beta, pad, seed=i + 1, naive=True))
    print(f"k2 small sweep {maj}{ta}{tb}: worst relerr {worst:.3e}")


def test_quick_returns_and_alpha_zero(u):
    M, N, K = 33, 17, 9
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, pad=(0, 0, 3), seed=4, sentinel=-5.0)
    # alpha == 0, beta == 1: untouched (sgemm_avx256.h:410)
    assert np.array_equal(gpu14(u, "auto", "R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 1.0, Cm, ldc), Cm)
    # alpha == 0: C <- beta*C on the M x N region of the GIVEN major (oracle = restated sgemm_cpu)
    for maj in ("R", "C"):
        got = gpu14(u, "auto", maj, "N", "N", M if maj == "R" else N, N if maj == "R" else M, K, 0.0, A, max(lda, M), B, max(ldb, N), 0.5, Cm, ldc)
        want = Cm.copy().reshape(M, ldc)
        want[:, :N] *= np.float32(0.5)
        assert np.array_equal(got.reshape(M, ldc), want)
    # K == 0 behaves like alpha == 0
    got = gpu14(u, "simt", "R", "N", "N", M, N, 0, 1.0, A, 1, B, N, 0.0, Cm, ldc)
    want = Cm.copy().reshape(M, ldc)
    want[:, :N] = 0
    assert np.array_equal(got.reshape(M, ldc), want)
    # M == 0 / N == 0
    assert np.array_equal(gpu14(u, "auto", "R", "N", "N", 0, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc), Cm)


def test_beta_zero_never_reads_c(u):
    """BLAS semantics shared by sgemm_avx / sgemm_c / sgemm_sse (sgemm_avx256.h:324-330): NaN in C is overwritten."""
    for mode, (M, N, K) in (("simt", (70, 50, 30)), ("3xtf32", (256, 256, 64))):
        A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=5)
        Cn = np.full_like(Cm, np.nan)
        got = gpu14(u, mode, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cn, ldc)
        want = oracle14("R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
        assert not np.isnan(got).any()
        assert O.relerr("R", M, N, want, got, ldc) <= TOL


def test_error_behaviour(u):
    A = np.ones(64 * 64, np.float32)
    Cm = np.full(64 * 64, 3.0, np.float32)
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda("R", "X", "N", 8, 8, 8, 1.0, A, 8, A, 8, 0.0, Cm, 8)
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda("R", "N", "N", 8, 8, 8, 1.0, A, 4, A, 8, 0.0, Cm, 8)       # lda < K
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_3xtf32("R", "N", "N", 32, 32, 32, 1.0, A, 33, A, 32, 0.0, Cm, 32)   # lda % 4 != 0: no silent fallback
    assert np.all(Cm == 3.0)
    assert u.last_error() is None
    # lower-case letters are accepted
    u.sgemm_cuda("r", "n", "t", 8, 8, 8, 1.0, A, 8, A, 8, 0.0, Cm, 8)
    assert np.all(Cm[:64] == 8.0)


K1_SHAPES = [(128, 128, 32), (256, 256, 64), (256, 512, 96), (384, 640, 200), (300, 200, 100), (1023, 1000, 1023),
             (129, 257, 33), (2048, 1024, 512)]


@pytest.mark.parametrize("cg", [2, 1])
@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_k1_all_transposes_vs_oracle(u, cg, ta, tb):
    """K1 (3xTF32 tcgen05) on tile-multiple and ragged shapes, ld multiples of 4, alpha/beta, both CTA-group modes."""
    u.set_k1_tuning(cta_group=cg)
    try:
        worst = 0.0
        for i, (M, N, K) in enumerate(K1_SHAPES):
            (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
            pa, pb, pc = (-ac) % 4, (-bc) % 4, (-N) % 4   # make every ld a multiple of 4 (TMA-eligible)
            for (alpha, beta, extra) in ((1.0, 0.0, 0), (1.5, 0.5, 4)):
                e = check_case(u, "3xtf32", "R", ta, tb, M, N, K, alpha, beta, (pa + extra, pb + extra, pc + extra), seed=10 + i)
                worst = max(worst, e)
            assert u.last_kernel() == "3xtf32"
        print(f"k1 cg={cg} {ta}{tb}: worst relerr {worst:.3e}")
    finally:
        u.set_k1_tuning(cta_group=0)


@pytest.mark.parametrize("cg", [2, 1])
def test_k1_ss_variant_and_rna_split_still_meet_the_gate(u, cg):
    """The round-1 SS kernel (both operands from shared memory) stays in the library for A/B runs and the RNA-split experiment:
    same shapes, same gate, and within round-off of the production TS kernel."""
    M, N, K = 384, 640, 200
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "T", "N", M, N, K, pad=(0, 0, 4), seed=21)
    want = oracle14("R", "T", "N", M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
    u.set_k1_tuning(cta_group=cg)
    try:
        ts = gpu14(u, "3xtf32", "R", "T", "N", M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        u.set_k1_variant(1)
        ss = gpu14(u, "3xtf32", "R", "T", "N", M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        worst = 0.0
        for i, (m, n, k) in enumerate(K1_SHAPES):
            for ta, tb in (("N", "N"), ("T", "T")):
                (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, m, n, k)
                worst = max(worst, check_case(u, "3xtf32", "R", ta, tb, m, n, k, 1.5, 0.5, ((-ac) % 4, (-bc) % 4, (-n) % 4 + 4), seed=40 + i))
        u.set_k1_variant(0)
        u.set_k1_tuning(split=1)           # round-to-nearest split: runs on the SS kernel whatever the variant
        rna = gpu14(u, "3xtf32", "R", "T", "N", M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
    finally:
        u.set_k1_variant(0)
        u.set_k1_tuning(split=0, cta_group=0)
    for name, got in (("TS", ts), ("SS", ss), ("SS, RNA split", rna)):
        e = O.relerr("R", M, N, want, got, ldc)
        print(f"cg={cg} {name}: relerr {e:.3e}")
        assert e <= TOL, (name, e)
    assert O.relerr("R", M, N, ts, ss, ldc) <= 2e-6
    print(f"SS variant cg={cg}: worst relerr over the K1 shapes {worst:.3e}")


def test_k1_promotion_interval_is_rounded_to_the_slice_count(u):
    """The TS kernel hands one 64-column slice to the epilogue every kc / NSL k-blocks, so kc is rounded up to a multiple of the
    slice count (4 for pair tiles, 2 for single-CTA tiles); any kc must give a correct result, and kc = 0 (never promote inside a
    tile) must show the truncating-accumulator bias the promotion exists for."""
    M, N, K = 512, 512, 4096
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=5)
    ref = (A.reshape(M, lda)[:, :K].astype(np.float64) @ B.reshape(K, ldb)[:, :N].astype(np.float64)).ravel()
    errs = {}
    try:
        for cg in (1, 2):
            for kc in (0, 1, 2, 3, 4, 6, 8):
                u.set_k1_tuning(kc_blocks=kc, cta_group=cg)
                got = gpu14(u, "3xtf32", "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
                errs[(cg, kc)] = float(np.linalg.norm(got.astype(np.float64) - ref) / np.linalg.norm(ref))
    finally:
        u.set_k1_tuning(kc_blocks=4, cta_group=0)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    for (cg, kc), e in errs.items():
        if kc:
            assert e <= TOL, (cg, kc, e)
    assert errs[(2, 0)] > errs[(2, 4)] * 5 and errs[(1, 0)] > errs[(1, 4)] * 5


def test_k1_column_major_and_zero_mean(u):
    for (maj, ta, tb) in (("C", "N", "N"), ("C", "T", "N"), ("C", "N", "T")):
        check_case(u, "3xtf32", maj, ta, tb, 384, 256, 160, 1.5, 0.5, (0, 0, 0), seed=3, lo=-0.5, hi=0.5)


def test_k1_vs_k2_agree(u):
    M, N, K = 512, 384, 777
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K + 3, seed=8)  # K+3: ld multiple of 4
    g1 = gpu14(u, "3xtf32", "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    g2 = gpu14(u, "simt", "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    assert O.relerr("R", M, N, g2, g1, ldc) <= TOL


def test_auto_dispatch_rule(u):
    """mode=auto: (1) K1 iff A,B 16-B aligned, lda/ldb % 4 == 0, M,N >= 128, K >= 32; (2) else, M,N >= 256 and K >= 64:
    repack the TMA-ineligible operand(s) to an aligned ld, then K1; (3) else K2."""
    for (M, N, K, pad, want, repacked) in ((256, 256, 64, (0, 0, 0), "3xtf32", False), (256, 256, 64, (1, 0, 0), "3xtf32", True),
                                           (256, 256, 64, (0, 3, 5), "3xtf32", True), (200, 256, 64, (1, 0, 0), "simt", False),
                                           (64, 256, 64, (0, 0, 0), "simt", False), (256, 256, 16, (0, 0, 0), "simt", False),
                                           (300, 257, 100, (1, 2, 3), "3xtf32", True)):
        check_case(u, "auto", "R", "N", "N", M, N, K, 1.5, 0.5, pad, seed=2)
        assert u.last_kernel() == want, (M, N, K, pad)
        assert u.last_repacked() == repacked, (M, N, K, pad)


def test_golden_fixtures(u):
    """GPU result vs outputs of the unmodified reference stored in tests/golden (all four CPU implementations)."""
    g = np.load(GOLDEN)
    cases = [ast.literal_eval(str(c)) for c in g["cases"]]
    for i, (maj, ta, tb, M, N, K, alpha, beta, pad, lo, hi) in enumerate(cases):
        A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=100 + i, lo=lo, hi=hi)
        (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
        for mode in ("simt", "auto"):
            got = gpu14(u, mode, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)
            for name in ("cpu", "c", "sse", "avx"):
                if f"{name}_{i}" in g:
                    e = O.relerr("R", cr, cc, g[f"{name}_{i}"], got, ldc)
                    assert e <= TOL, (i, mode, name, e)


def test_config1_1024_cube_vs_reference_avx(u):
    """BASELINE config 1: row-major NN 1024^3 alpha=1 beta=0, oracle = sgemm_avx (live _ref when shipped,
    else the bit-identical 35-band restatement).  The reference's own cmp_results line is printed too."""
    M = N = K = 1024
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=1)
    r = O.ref()
    if r is not None:
        want = O.run14(r.ref_sgemm_avx, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    else:
        want = oracle14("R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    for mode in ("3xtf32", "simt"):
        got = gpu14(u, mode, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
        e = O.relerr("R", M, N, want, got, ldc)
        out = np.zeros(4)
        verdict = O.oracle().oracle_cmp_results(M, N, want, got, ldc, out)
        print(f"c1 {mode}: relerr {e:.3e} ({bounds(K)}); cmp_results stdErr/stdRef={out[0] / out[1]:.3e} maxErr={out[2]:.3e} verdict={verdict}")
        assert e <= TOL


@pytest.mark.parametrize("ta,tb", [("N", "T"), ("T", "N"), ("T", "T"), ("N", "N")])
def test_config3_ragged_transposed(u, ta, tb):
    """BASELINE config 3: 4095 x 3001 x 2047, alpha=1.5 beta=0.5, (i) ld padded to multiples of 4 -> K1,
    (ii) odd padding -> auto repacks to an aligned ld and runs K1, forced simt runs K2; padding untouched.  Oracle: reference sgemm_sse when _ref ships, else 35-band."""
    M, N, K = 4095, 3001, 2047
    r = O.ref()
    (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
    odd = lambda w, p: p if (w + p) % 4 else p + 1   # a pad that leaves the leading dimension NOT a multiple of 4
    for label, pad, mode, want_kernel in (("ld%4==0", ((-ac) % 4, (-bc) % 4, (-N) % 4), "auto", "3xtf32"),
                                          ("odd ld, auto (repack + K1)", (odd(ac, 5), odd(bc, 3), odd(N, 7)), "auto", "3xtf32"),
                                          ("odd ld, forced K2", (odd(ac, 5), odd(bc, 3), odd(N, 7)), "simt", "simt")):
        A, lda, B, ldb, Cm, ldc = O.make_problem("R", ta, tb, M, N, K, pad=pad, seed=33, sentinel=-9.0)
        if r is not None:
            want = O.run14(r.ref_sgemm_sse, "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        else:
            want = oracle14("R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        got = gpu14(u, mode, "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        assert u.last_kernel() == want_kernel
        assert u.last_repacked() == label.startswith("odd ld, auto")
        e = O.relerr("R", M, N, want, got, ldc)
        print(f"c3 {ta}{tb} {label} -> {want_kernel}: relerr {e:.3e} ({bounds(K)})")
        assert e <= TOL
        assert np.array_equal(got.reshape(M, ldc)[:, N:], Cm.reshape(M, ldc)[:, N:])


def sampled_rows_check(u, mode, M, N, K, rows, lo, hi, seedA=1, seedB=2):
    """Full-size NN problem generated ON THE DEVICE from the shared RNG; the oracle recomputes only `rows`."""
    dA = u.DeviceBuffer(M * K).fill_uniform(seedA, lo, hi)
    dB = u.DeviceBuffer(K * N).fill_uniform(seedB, lo, hi)
    dC = u.DeviceBuffer(M * N)
    u.sgemm_cuda_dev(mode, None, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
    u.sync()
    B = O.fill_uniform(K * N, seedB, lo, hi)
    worst = 0.0
    for r0, nr in rows:
        full = O.fill_uniform((r0 + nr) * K, seedA, lo, hi) if (r0 + nr) * K <= (1 << 27) else None
        if full is not None:
            A = full[r0 * K:(r0 + nr) * K].copy()
        else:  # regenerate just the slab: element i of the stream is independent of the others
            import ugemm_b200 as mod
            tmp = mod.DeviceBuffer(nr * K)
            # stream offset: fill with the same seed but shifted base is not exposed; download from dA instead
            A = dA.download(nr * K, offset=r0 * K)
            tmp.free()
        C0 = np.zeros(nr * N, np.float32)
        want = oracle14("R", "N", "N", nr, N, K, 1.0, A, K, B, N, 0.0, C0, N)
        got = dC.download(nr * N, offset=r0 * N)
        worst = max(worst, O.relerr("R", nr, N, want, got, N))
    for d in (dA, dB, dC):
        d.free()
    return worst


@pytest.mark.parametrize("lo,hi", [(0.0, 1.0), (-0.5, 0.5)])
def test_config2_8192_cube_sampled(u, lo, hi):
    """BASELINE config 2 (8192^3 NN): sampled 16-row slabs vs the oracle, both input distributions, both kernels."""
    M = N = K = 8192
    rows = [(0, 16), (4095, 16), (8176, 16)]
    for mode in ("3xtf32", "simt"):
        e = sampled_rows_check(u, mode, M, N, K, rows, lo, hi)
        print(f"c2 {mode} U[{lo},{hi}): sampled relerr {e:.3e} ({bounds(K)})")
        assert e <= TOL


def test_config4_tall_skinny_sampled(u):
    """BASELINE config 4: im2col-shaped 200704 x 256 x 1152 NN."""
    M, N, K = 200704, 256, 1152
    rows = [(0, 64), (100000, 64), (200640, 64)]
    for mode in ("3xtf32", "simt"):
        e = sampled_rows_check(u, mode, M, N, K, rows, 0.0, 1.0)
        print(f"c4 {mode}: sampled relerr {e:.3e} ({bounds(K)})")
        assert e <= TOL


def test_non_finite_inputs(u):
    """Inf / NaN in A or B (include/ugemm_cuda.h, "Non-finite inputs").  K2 propagates them like the reference's loops: the same
    entries are NaN, +Inf, -Inf.  K1 marks the same entries non-finite (an Inf of the reference may be a NaN there) and leaves
    every entry with finite inputs within the gate."""
    M, N, K = 256, 256, 128
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=77, lo=-0.5, hi=0.5)
    A = A.copy(); B = B.copy()
    A[3 * lda + 5] = np.inf           # row 3 of C
    A[100 * lda + 64] = -np.inf       # row 100
    A[200 * lda + 127] = np.nan       # row 200
    B[17 * ldb + 9] = np.inf          # column 9
    B[40 * ldb + 130] = 1.0           # exactly representable in TF32: b_small == 0 -> K1's Inf * 0
    with np.errstate(invalid="ignore", over="ignore"):
        want = oracle14("R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc).reshape(M, ldc)[:, :N]
    fin = np.isfinite(want)
    assert not fin[3].any() and not fin[100].any() and not fin[200].any() and not fin[:, 9].any() and fin.sum() == (M - 3) * (N - 1)
    for mode in ("simt", "3xtf32"):
        got = gpu14(u, mode, "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc).reshape(M, ldc)[:, :N]
        assert u.last_kernel() == mode
        assert np.array_equal(np.isfinite(got), fin), mode
        e = np.linalg.norm(got[fin].astype(np.float64) - want[fin]) / np.linalg.norm(want[fin].astype(np.float64))
        assert e <= TOL, (mode, e)
        if mode == "simt":
            assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.sign(got[np.isinf(got)]), np.sign(want[np.isinf(want)]))
        else:   # K1: at least the entries of row 3 whose other factors have a non-zero small part keep their Inf
            assert np.isinf(got[3]).sum() > 0 and np.isnan(got[200]).all()


def test_second_init_on_another_device_is_an_error(u):
    """One backend, one device: sgemm_cuda_init on a different ordinal while initialised must fail loudly."""
    u.sgemm_cuda_init(0)                      # same device: fine
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_init(1)
    u.sgemm_cuda_init(0)
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", 64, 64, 64, seed=5)
    gpu14(u, "auto", "R", "N", "N", 64, 64, 64, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)   # the backend is still usable


def test_ablation_flags_need_an_opt_in():
    """UGEMM_K1_FLAGS with result-corrupting bits is rejected at init unless UGEMM_K1_ABLATION=1 (fresh process each)."""
    import subprocess
    import sys
    code = "import ugemm_b200 as u\ntry:\n    u.sgemm_cuda_init(0); print('INIT-OK')\nexcept u.UgemmCudaError as e:\n    print('INIT-ERR', e)\n"
    env = dict(os.environ, UGEMM_K1_FLAGS="9")
    env.pop("UGEMM_K1_ABLATION", None)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT).stdout
    assert "INIT-ERR" in out and "ablation" in out
    out = subprocess.run([sys.executable, "-c", code], env=dict(env, UGEMM_K1_ABLATION="1"), capture_output=True, text=True, cwd=ROOT).stdout
    assert "INIT-OK" in out
    out = subprocess.run([sys.executable, "-c", code], env=dict(env, UGEMM_K1_FLAGS="2049"), capture_output=True, text=True, cwd=ROOT).stdout
    assert "INIT-OK" in out                    # tuning bits (collector, stream-K off) are always allowed


def test_config5_32768_sampled(u):
    """BASELINE config 5 (32768^3 NN, alpha=1, beta=0, device-generated operands): 32 sampled rows -- 4 slabs of 8, one of them
    straddling the M/2 block boundary of every grid -- against the reference's sgemm_avx (sgemm_avx256.h:392) on the same rows
    regenerated from the shared RNG.  Through sgemm_cuda_dev on one GPU (K1) and through sgemm_cuda_mgpu on the largest
    BASELINE grid the box holds (1x1, 2x1, 2x2, 2x4)."""
    r = O.ref()
    M = N = K = 32768
    lo, hi = -0.5, 0.5
    slabs = [(0, 8), (16380, 8), (24571, 8), (32760, 8)]
    dA = u.DeviceBuffer(M * K).fill_uniform(1, lo, hi)
    dB = u.DeviceBuffer(K * N).fill_uniform(2, lo, hi)
    dC = u.DeviceBuffer(M * N)
    B = O.fill_uniform_at(K * N, 2, 0, lo, hi)                      # 4.3 GB, regenerated on the host (never downloaded)
    wants = []
    for r0, nr in slabs:
        A = O.fill_uniform_at(nr * K, 1, r0 * K, lo, hi)
        C0 = np.zeros(nr * N, np.float32)
        if r is not None:
            wants.append(O.run14(r.ref_sgemm_avx, "R", "N", "N", nr, N, K, 1.0, A, K, B, N, 0.0, C0, N))
        else:   # GPU box without a prebuilt oracle/_ref: the bit-identical restatement (tests/test_oracle.py pins it to sgemm_avx)
            wants.append(O.run14(O.oracle().oracle_sgemm_banded, "R", "N", "N", nr, N, K, 1.0, A, K, B, N, 0.0, C0, N, threads=1))
    del B

    def worst_of(tag):
        worst = 0.0
        for (r0, nr), want in zip(slabs, wants):
            got = dC.download(nr * N, offset=r0 * N)
            worst = max(worst, O.relerr("R", nr, N, want, got, N))
        print(f"c5 {tag}: sampled relerr {worst:.3e} over 32 rows ({bounds(K)})")
        return worst

    try:
        u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
        u.sync()
        assert u.last_kernel() == "3xtf32"
        assert worst_of("sgemm_cuda_dev 1 GPU") <= TOL
        n = min(u.visible_gpus(), 8)
        pr, pc = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}[8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1]
        dC.upload(np.full(1 << 20, np.nan, np.float32))              # poison the head of C so a skipped run cannot pass
        u.sgemm_cuda_mgpu_init(pr * pc)
        try:
            u.sgemm_cuda_mgpu("R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N, pr, pc, 1)
        finally:
            u.sgemm_cuda_mgpu_finish()
        assert worst_of(f"sgemm_cuda_mgpu {pr}x{pc}") <= TOL
    finally:
        for d in (dA, dB, dC):
            d.free()


def test_full_size_linearity_property(u):
    """Size-independent property at 4096^3: GEMM(A, B1 + B2) == GEMM(A, B1) + GEMM(A, B2) (beta=1 accumulate path)."""
    M = N = K = 4096
    dA = u.DeviceBuffer(M * K).fill_uniform(1, -0.5, 0.5)
    b1 = u.fill_uniform_host(K * N, 2, -0.5, 0.5)
    b2 = u.fill_uniform_host(K * N, 3, -0.5, 0.5)
    dB1, dB2, dB12 = u.DeviceBuffer(K * N).upload(b1), u.DeviceBuffer(K * N).upload(b2), u.DeviceBuffer(K * N).upload(b1 + b2)
    dC, dD = u.DeviceBuffer(M * N), u.DeviceBuffer(M * N)
    u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, K, dB1, N, 0.0, dC, N)
    u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, K, dB2, N, 1.0, dC, N)
    u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, K, dB12, N, 0.0, dD, N)
    u.sync()
    c, d = dC.download(), dD.download()
    e = O.relerr("R", M, N, d, c, N)
    print(f"linearity 4096^3: {e:.3e}")
    assert e <= TOL
    for x in (dA, dB1, dB2, dB12, dC, dD):
        x.free()


@pytest.mark.parametrize("case", [
    ("R", "N", "N", 128, 361, 1152, 11, (0, 0, 0), "simt"),      # the reference's default check_sgemm problem: 11 stacked instances
    ("R", "N", "N", 128, 360, 1152, 11, (0, 0, 0), "3xtf32"),    # same with a TMA-eligible ldb -> one K1 launch over 11 x 3 tiles
    ("R", "T", "N", 256, 384, 200, 5, (0, 0, 4), "3xtf32"),
    ("R", "N", "T", 300, 260, 100, 3, (4, 4, 0), "3xtf32"),
    ("C", "N", "N", 130, 140, 70, 4, (2, 0, 1), "simt"),
])
def test_batched_stacked_instances(u, case):
    """sgemm_cuda_batched == the reference's loop over stacked instances (check_sgemm.c:111-124), one launch."""
    maj, ta, tb, M, N, K, batch, pad, want_kernel = case
    (ar, ac), (br, bc), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
    lda, ldb, ldc = ac + pad[0], bc + pad[1], cc + pad[2]
    sA, sB, sC = ar * lda, br * ldb, cr * ldc
    A = O.fill_uniform(batch * sA, 901, -0.5, 0.5)
    B = O.fill_uniform(batch * sB, 902, -0.5, 0.5)
    C0 = O.fill_uniform(batch * sC, 903, -0.5, 0.5)
    got = C0.copy()
    u.sgemm_cuda_batched(maj, ta, tb, M, N, K, 1.5, A, lda, sA, B, ldb, sB, 0.5, got, ldc, sC, batch)
    assert u.last_kernel() == want_kernel
    for b in range(batch):
        want = oracle14(maj, ta, tb, M, N, K, 1.5, A[b * sA:(b + 1) * sA], lda, B[b * sB:(b + 1) * sB], ldb, 0.5,
                        C0[b * sC:(b + 1) * sC], ldc)
        e = O.relerr("R", cr, cc, want, got[b * sC:(b + 1) * sC], ldc)
        assert e <= TOL, (b, e)
        if pad[2]:
            assert np.array_equal(got[b * sC:(b + 1) * sC].reshape(cr, ldc)[:, cc:], C0[b * sC:(b + 1) * sC].reshape(cr, ldc)[:, cc:])


def test_batched_k1_with_an_unaligned_instance_stride_of_c(u):
    """ldc % 4 == 0 but strideC % 4 != 0: rows of instances 1.. do not start on 16-byte boundaries, so K1 must fall back to scalar
    accesses to C (no TMA boxes, no 128-bit loads or stores) -- and still match the per-instance oracle loop; gaps untouched."""
    M, N, K, batch = 256, 384, 200, 3
    lda, ldb, ldc = K, N, N
    sA, sB, sC = M * lda, K * ldb, M * ldc + 2
    A = O.fill_uniform(batch * sA, 911, -0.5, 0.5)
    B = O.fill_uniform(batch * sB, 912, -0.5, 0.5)
    C0 = O.fill_uniform(batch * sC, 913, -0.5, 0.5)
    got = C0.copy()
    u.sgemm_cuda_batched("R", "N", "N", M, N, K, 1.5, A, lda, sA, B, ldb, sB, 0.5, got, ldc, sC, batch)
    assert u.last_kernel() == "3xtf32"
    for b in range(batch):
        want = oracle14("R", "N", "N", M, N, K, 1.5, A[b * sA:(b + 1) * sA], lda, B[b * sB:(b + 1) * sB], ldb, 0.5, C0[b * sC:b * sC + M * ldc], ldc)
        e = O.relerr("R", M, N, want, got[b * sC:b * sC + M * ldc], ldc)
        assert e <= TOL, (b, e)
        assert np.array_equal(got[b * sC + M * ldc:(b + 1) * sC], C0[b * sC + M * ldc:(b + 1) * sC]), "the gap between instances was written"


def test_batched_host_scale_only_and_stride_checks(u):
    """sgemm_cuda_batched (host pointers): K == 0 / alpha == 0 never read A or B (NULL is legal, transposed operands included);
    bad strides are an error that leaves C untouched."""
    M, N, batch, ldc = 37, 29, 3, 32
    sC = M * ldc
    C0 = O.fill_uniform(batch * sC, 911, -0.5, 0.5)
    for K, alpha, ta in ((0, 1.5, "T"), (16, 0.0, "N"), (0, 0.0, "T")):
        got = C0.copy()
        u.sgemm_cuda_batched("R", ta, "N", M, N, K, alpha, None, max(M, K, 1), 0, None, N, 0, 0.5, got, ldc, sC, batch)
        want = C0.copy().reshape(batch, M, ldc)
        want[:, :, :N] *= np.float32(0.5)
        assert np.array_equal(got.reshape(batch, M, ldc), want)
    A = O.fill_uniform(batch * M * 16, 912)
    B = O.fill_uniform(batch * 16 * N, 913)
    for strides in ((-1, 16 * N, sC), (M * 16, -5, sC), (M * 16, 16 * N, sC - 40)):
        got = C0.copy()
        with pytest.raises(u.UgemmCudaError):
            u.sgemm_cuda_batched("R", "N", "N", M, N, 16, 1.0, A, 16, strides[0], B, N, strides[1], 0.0, got, ldc, strides[2], batch)
        assert np.array_equal(got, C0)
    got = C0.copy()   # K == 0 and beta == 1: quick return
    u.sgemm_cuda_batched("R", "N", "N", M, N, 0, 1.0, None, 1, 0, None, N, 0, 1.0, got, ldc, sC, batch)
    assert np.array_equal(got, C0)


def test_pipelined_host_path_uses_one_kernel_for_every_panel(u):
    """M = 4097 on the panel-pipelined host path (>= 64 MB in flight): the one-row tail joins the previous panel and every panel runs
    the kernel chosen for the whole problem (K1).  (Not bit-identical to the one-shot launch: the stream-K tail cuts
    different tiles along K in the two launches, which regroups -- not changes -- the fp32 partial sums.)"""
    M, N, K = 4097, 2048, 2048
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, seed=321, lo=-0.5, hi=0.5)
    piped = gpu14(u, "auto", "R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
    assert u.last_kernel() == "3xtf32"
    dA, dB, dC = u.DeviceBuffer(A.size).upload(A), u.DeviceBuffer(B.size).upload(B), u.DeviceBuffer(Cm.size)
    u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", M, N, K, 1.0, dA, lda, dB, ldb, 0.0, dC, ldc)
    u.sync()
    whole = dC.download()
    rows = np.r_[0:8, 2040:2056, 4088:4097]
    want = (A.reshape(M, K)[rows].astype(np.float64) @ B.reshape(K, N).astype(np.float64))
    for got in (piped, whole):
        g = got.reshape(M, N)[rows]
        assert np.linalg.norm(g - want) / np.linalg.norm(want) <= TOL
    assert O.relerr("R", M, N, whole, piped, N) <= 2e-6
    for d in (dA, dB, dC):
        d.free()


@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_k2_narrow_n_tiles(u, ta, tb):
    """Tall-skinny shapes (N <= 32, M >= 256) take K2's 256x16 / 256x32 tiles; ragged edges, odd ld, both majors."""
    for i, (M, N, K) in enumerate([(300, 16, 70), (257, 7, 33), (1000, 32, 129), (512, 24, 64), (4096, 1, 100), (777, 31, 1)]):
        for maj in ("R", "C"):
            pad = (0, 0, 0) if i % 2 == 0 else (3, 1, 5)
            mm, nn = (M, N) if maj == "R" else (N, M)   # column-major swaps the roles, so keep the skinny side on N after the swap
            check_case(u, "simt", maj, ta, tb, mm, nn, K, 1.5, 0.5, pad, seed=60 + i, naive=True)
    # the memory-bound corner at size: 200704 x 16 x 1152 on K2's narrow tiles (forced: since round 2 the auto rule sends this
    # shape to K1, which skips the accumulator slices without columns and is faster from a narrow side of 8 when K >= 512) ...
    e = sampled_rows_check(u, "simt", 200704, 16, 1152, [(0, 32), (200672, 32)], 0.0, 1.0)
    assert u.last_kernel() == "simt"
    assert e <= TOL
    # ... and through auto, on K1
    e = sampled_rows_check(u, "auto", 200704, 16, 1152, [(0, 32), (200672, 32)], 0.0, 1.0)
    assert u.last_kernel() == "3xtf32"
    assert e <= TOL


@pytest.mark.parametrize("ta", ["N", "T"])
def test_k2_big_tiles_interior_and_edges(u, ta):
    """K2 with >= one 128x128 tile per SM: whole-tile shapes (the unguarded interior loop), K tails that end inside a 128-bit
    quad / inside a k-tile, and ragged M / N whose last tile row and column take the guarded loop."""
    for i, (M, N, K) in enumerate(((1664, 1536, 64), (1664, 1536, 100), (1664, 1536, 19), (1701, 1541, 70), (2048, 1280, 1), (1536, 1700, 33))):
        for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
            (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, "N", M, N, K)
            pad = ((-ac) % 4, (-bc) % 4, 3)      # lda, ldb multiples of 4 (16-byte copies), ldc odd
            check_case(u, "simt", "R", ta, "N", M, N, K, alpha, beta, pad, seed=80 + i)
            assert u.last_kernel() == "simt"
    # column-major maps onto the same kernel with the operands swapped
    check_case(u, "simt", "C", "N", ta, 1536, 1664, 50, 1.5, 0.5, (0, 0, 0), seed=90)


def test_k1_stream_k_tail_is_deterministic_and_exact_enough(u):
    """Shapes whose tile count is not a multiple of the SM-pair count finish their last round by stream-K (tail tiles cut along K,
    partial sums added in a fixed order by a second kernel): results must be reproducible bit for bit and meet the gate, for
    ragged edges, beta != 0 and the bias + LeakyReLU epilogue's plain path."""
    for i, (M, N, K, ta, tb) in enumerate(((1100, 900, 1024, "N", "N"), (4095, 3001, 2047, "N", "T"), (2048, 2048, 2048, "N", "N"),
                                           (768, 640, 4096, "T", "N"), (4096, 4096, 512, "N", "N"))):
        (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
        pad = ((-ac) % 4, (-bc) % 4, (-N) % 4 + 4)
        A, lda, B, ldb, Cm, ldc = O.make_problem("R", ta, tb, M, N, K, pad=pad, seed=60 + i, sentinel=-77.0)
        g1 = gpu14(u, "3xtf32", "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        g2 = gpu14(u, "3xtf32", "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
        assert np.array_equal(g1, g2), "stream-K result changed between two identical calls"
        assert np.array_equal(g1.reshape(M, ldc)[:, N:], Cm.reshape(M, ldc)[:, N:]), "ld padding of C was written"
        if M * N * K <= 2.2e9:
            want = oracle14("R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, 0.5, Cm, ldc)
            assert O.relerr("R", M, N, want, g1, ldc) <= TOL
        else:   # sampled rows against fp64
            opA = A.reshape(ar, lda)[:, :ac] if ta == "N" else A.reshape(ar, lda)[:, :ac].T
            opB = B.reshape(br, ldb)[:, :bc] if tb == "N" else B.reshape(br, ldb)[:, :bc].T
            rows = np.linspace(0, M - 1, 24).astype(int)
            ref = 1.5 * (opA[rows].astype(np.float64) @ opB.astype(np.float64)) + 0.5 * Cm.reshape(M, ldc)[rows, :N]
            got = g1.reshape(M, ldc)[rows, :N]
            assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= TOL


@pytest.mark.parametrize("cg", [2, 1])
def test_k1_old_c_through_tma_and_through_registers(u, cg):
    """beta != 0 on K1: a C the TMA unit can address (16-byte aligned, ldc % 4 == 0) is fetched box by box into the epilogue warps'
    staging boxes and added between two promotions at a fixed hand-over number; any other C is loaded into the registers.  Both
    against the oracle (the old C is 2 % of the result at the shortest K here: a missing or misplaced 32 x 32 box is 1000 x the
    gate), reproducible bit for bit, ld padding untouched.  Shapes: boxes clipped at both edges, K so short that the hand-overs
    run out before a warp's four boxes are in, and -- on 8 SMs -- many tiles per CTA, so that every staging box alternates
    between C loads and tile stores."""
    u.set_k1_tuning(cta_group=cg)
    try:
        worst = 0.0
        for sms in (0, 8):
            u.set_sm_limit(sms)
            for i, (M, N, K, ta, tb) in enumerate(((700, 500, 96, "N", "T"), (300, 260, 100, "T", "N"), (1100, 900, 1024, "N", "N"), (513, 1030, 160, "T", "T"))):
                if sms and K > 512:
                    continue
                (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
                for padc, beta in ((0, 0.5), (4, 1.0), (3, -2.0), (1, 0.5)):
                    pad = ((-ac) % 4, (-bc) % 4, (-N) % 4 + padc)
                    A, lda, B, ldb, Cm, ldc = O.make_problem("R", ta, tb, M, N, K, pad=pad, seed=80 + i, sentinel=-77.0)
                    g1 = gpu14(u, "3xtf32", "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, beta, Cm, ldc)
                    assert u.last_kernel() == "3xtf32"
                    g2 = gpu14(u, "3xtf32", "R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, beta, Cm, ldc)
                    assert np.array_equal(g1, g2), f"result changed between two identical calls ({M}x{N}x{K} ldc={ldc} beta={beta} sms={sms})"
                    assert np.array_equal(g1.reshape(M, ldc)[:, N:], Cm.reshape(M, ldc)[:, N:]), "ld padding of C was written"
                    want = oracle14("R", ta, tb, M, N, K, 1.5, A, lda, B, ldb, beta, Cm, ldc)
                    e = O.relerr("R", M, N, want, g1, ldc)
                    assert e <= TOL, f"{ta}{tb} {M}x{N}x{K} ldc={ldc} beta={beta} sms={sms}: relerr {e:.3e}"
                    worst = max(worst, e)
        print(f"k1 old C, cg={cg}: worst relerr {worst:.3e}")
    finally:
        u.set_sm_limit(0)
        u.set_k1_tuning(cta_group=0)


def test_auto_dispatch_skinny_but_large_goes_to_k1(u):
    """One side >= 128, the other >= 48 (>= 8 with K >= 512), M*N*K >= 2^26: K1 on a zero-filled tile beats K2 (dispatch rule)."""
    for (M, N, K, want) in ((4096, 64, 512, "3xtf32"), (64, 4096, 512, "3xtf32"), (8192, 48, 256, "3xtf32"), (4096, 32, 512, "3xtf32"),
                            (16384, 8, 1024, "3xtf32"), (8, 16384, 1024, "3xtf32"), (8192, 32, 256, "simt"), (16384, 4, 2048, "simt"),
                            (4096, 64, 128, "simt"), (100, 100, 8192, "simt")):
        for ta, tb in (("N", "N"), ("T", "T")):
            (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
            check_case(u, "auto", "R", ta, tb, M, N, K, 1.5, 0.5, ((-ac) % 4, (-bc) % 4, 1), seed=33)
            assert u.last_kernel() == want, (M, N, K, ta, tb)
