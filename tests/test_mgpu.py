"""Single-process multi-GPU entry points (sgemm_cuda_mgpu_*, -m gpu): the pr x pc block grid with the pipelined panel
relay against the oracle on identical inputs, same gate as the single-GPU parity tests (relerr <= 1e-5, ld padding
bit-identical).  Grids larger than the box are skipped (the 1 x 1 grid and the slab arithmetic run on any box)."""
import numpy as np
import pytest

import _oracle as O
from test_parity_gpu import TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg(u):
    n = min(u.visible_gpus(), 8)
    assert n >= 1
    u.sgemm_cuda_mgpu_init(n)
    assert u.sgemm_cuda_mgpu_count() == n
    yield n
    u.sgemm_cuda_mgpu_finish()
    assert u.sgemm_cuda_mgpu_count() == 0


def run_case(u, maj, ta, tb, M, N, K, alpha, beta, pad, pr, pc, overlap, seed=1, lo=-0.5, hi=0.5):
    A, lda, B, ldb, Cm, ldc = O.make_problem(maj, ta, tb, M, N, K, pad=pad, seed=seed, lo=lo, hi=hi, sentinel=-77.0)
    got = Cm.copy()
    t = u.sgemm_cuda_mgpu(maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, got, ldc, pr, pc, overlap)
    want = O.run14(O.oracle().oracle_sgemm_banded, maj, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=8)
    (_, _), (_, _), (cr, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
    e = O.relerr("R", cr, cc, want, got, ldc)
    assert e <= TOL, f"mgpu {pr}x{pc} ov={overlap} {maj}{ta}{tb} {M}x{N}x{K} a={alpha} b={beta} pad={pad}: relerr {e:.3e}"
    if pad[2]:
        assert np.array_equal(got.reshape(cr, ldc)[:, cc:], Cm.reshape(cr, ldc)[:, cc:]), "ld padding of C was written"
    return e, t


GRIDS = [(1, 1), (2, 1), (1, 2), (2, 2), (2, 4), (4, 2)]


@pytest.mark.parametrize("pr,pc", GRIDS)
def test_grid_all_transposes(u, mg, pr, pc):
    if pr * pc > mg:
        pytest.skip(f"{pr}x{pc} grid needs {pr * pc} GPUs, box has {mg}")
    worst = 0.0
    i = 0
    for maj in ("R", "C"):
        for ta in ("N", "T"):
            for tb in ("N", "T"):
                i += 1
                # ragged blocks (M, N not multiples of the grid or of a tile), K cut into 3 slabs when overlapped
                e, t = run_case(u, maj, ta, tb, 1001, 778, 8300, 1.5, 0.5, (4, 8, 4) if i % 2 else (1, 3, 5), pr, pc, 2 if i % 3 else 0, seed=i)
                worst = max(worst, e)
    print(f"mgpu {pr}x{pc}: worst relerr {worst:.3e}, last timings ms {tuple(round(x, 3) for x in t)}")


def test_small_and_degenerate_blocks(u, mg):
    """Problems smaller than the grid (empty blocks), one-row / one-column blocks, K below one k-block."""
    pr, pc = (2, 1) if mg < 4 else (2, 2)
    if mg < 2:
        pr, pc = 1, 1
    for (M, N, K) in ((1, 1, 1), (3, 5, 2), (4, 300, 17), (300, 2, 33), (129, 130, 31)):
        run_case(u, "R", "N", "N", M, N, K, 1.0, 0.0, (0, 0, 0), pr, pc, 1, seed=M + N)
        run_case(u, "C", "T", "N", M, N, K, -1.0, 2.0, (2, 1, 3), pr, pc, 1, seed=M + N + 1)


def test_quick_returns_scale_only_and_nan(u, mg):
    pr, pc = (2, 1) if mg >= 2 else (1, 1)
    M, N, K = 257, 130, 64
    A, lda, B, ldb, Cm, ldc = O.make_problem("R", "N", "N", M, N, K, pad=(0, 0, 2), seed=5, sentinel=-3.0)
    # alpha == 0, beta == 1: untouched (sgemm_avx256.h:410)
    got = Cm.copy()
    u.sgemm_cuda_mgpu("R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 1.0, got, ldc, pr, pc)
    assert np.array_equal(got, Cm)
    # alpha == 0: C <- beta * C on the M x N region only
    got = Cm.copy()
    u.sgemm_cuda_mgpu("R", "N", "N", M, N, K, 0.0, A, lda, B, ldb, 0.5, got, ldc, pr, pc)
    want = Cm.copy().reshape(M, ldc)
    want[:, :N] *= np.float32(0.5)
    assert np.array_equal(got.reshape(M, ldc), want)
    # K == 0, beta == 0: zeros, padding untouched
    got = Cm.copy()
    u.sgemm_cuda_mgpu("R", "N", "N", M, N, 0, 1.0, A, 1, B, N, 0.0, got, ldc, pr, pc)
    want = Cm.copy().reshape(M, ldc)
    want[:, :N] = 0
    assert np.array_equal(got.reshape(M, ldc), want)
    # beta == 0 never reads C: NaN in C must not propagate
    bad = Cm.copy()
    bad.reshape(M, ldc)[:, :N] = np.nan
    u.sgemm_cuda_mgpu("R", "N", "N", M, N, K, 1.0, A, lda, B, ldb, 0.0, bad, ldc, pr, pc)
    assert np.isfinite(bad.reshape(M, ldc)[:, :N]).all()
    # M == 0
    got = Cm.copy()
    u.sgemm_cuda_mgpu("R", "N", "N", 0, N, K, 1.0, A, lda, B, ldb, 0.0, got, ldc, pr, pc)
    assert np.array_equal(got, Cm)


def test_device_resident_operands_and_timings(u, mg):
    """Operands already on GPU 0 (the placement the C harness uses for config 5): relay starts from device memory."""
    pr, pc = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}.get(mg, (1, mg))
    M, N, K = 2048, 1536, 4608
    dA, dB, dC = u.DeviceBuffer(M * K), u.DeviceBuffer(K * N), u.DeviceBuffer(M * N)
    try:
        dA.fill_uniform(11, -0.5, 0.5)
        dB.fill_uniform(12, -0.5, 0.5)
        dC.fill_uniform(13, 0.0, 1.0)
        A, B, C0 = dA.download(), dB.download(), dC.download()
        for overlap in (0, 1):
            dC.upload(C0)
            t = u.sgemm_cuda_mgpu("R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 1.0, dC, N, pr, pc, overlap)
            got = dC.download()
            want = O.run14(O.oracle().oracle_sgemm_banded, "R", "N", "N", M, N, K, 1.0, A, K, B, N, 1.0, C0, N, threads=8)
            assert O.relerr("R", M, N, want, got, N) <= TOL
            assert t[1] > 0 and t[3] > 0 and t[0] >= t[1] * 0.5
            print(f"mgpu {pr}x{pc} device-resident overlap={overlap}: wall {t[0]:.3f} span {t[1]:.3f} dist {t[2]:.3f} gemm {t[3]:.3f} ms")
    finally:
        dA.free(); dB.free(); dC.free()


def test_errors(u, mg):
    A = np.ones(64, np.float32)
    Cm = np.full(64, 3.0, np.float32)
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_mgpu("R", "N", "N", 8, 8, 8, 1.0, A, 8, A, 8, 0.0, Cm, 8, mg + 1, 1)      # grid larger than the box
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_mgpu("R", "N", "N", 8, 8, 8, 1.0, A, 4, A, 8, 0.0, Cm, 8, 1, 1)           # lda < K
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_mgpu("R", "Q", "N", 8, 8, 8, 1.0, A, 8, A, 8, 0.0, Cm, 8, 1, 1)
    assert np.all(Cm == 3.0)
    assert u.last_error() is None
