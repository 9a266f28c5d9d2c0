"""Seeded random parity sweep (-m gpu): shapes that straddle every threshold of the dispatch rule, all eight
major/transpose cases, tight / 4-multiple / odd leading dimensions, the alpha/beta quick-return corners, and
device pointers that are only 4-byte aligned.  Same gate as tests/test_parity_gpu.py (normwise relerr <= 1e-5
against the oracle, ld padding bit-identical); the reference's harness draws its cases the same way, from
generated data with no stored outputs (check_sgemm.c:47-54,145-257)."""
import numpy as np
import pytest

import _oracle as O
from test_parity_gpu import TOL, check_case

pytestmark = pytest.mark.gpu

SCALARS = [(1.0, 0.0), (1.5, 0.5), (1.0, 1.0), (-1.0, 2.0), (0.0, 0.5), (2.0, 0.0), (0.0, 1.0), (0.0, 0.0)]
# dimensions around the rule's thresholds (48, 128, 256), the tile sizes (64, 128, 256) and the k-block (32)
EDGES = [1, 2, 5, 31, 32, 33, 47, 48, 63, 64, 65, 127, 128, 129, 191, 255, 256, 257, 300, 383, 385, 511, 513, 640]
K_EDGES = [1, 3, 8, 16, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 255, 257, 384, 511]


def _draw(rng, i):
    maj = "RC"[int(rng.integers(2))]
    ta, tb = "NT"[int(rng.integers(2))], "NT"[int(rng.integers(2))]
    M, N = int(rng.choice(EDGES)), int(rng.choice(EDGES))
    K = int(rng.choice(K_EDGES))
    alpha, beta = SCALARS[i % len(SCALARS)]
    kind = i % 3
    if kind == 0:
        pad = (0, 0, 0)
    elif kind == 1:                               # every ld a multiple of 4 -> TMA-eligible when big enough
        (_, ac), (_, bc), (_, cc) = O.stored_shapes(maj, ta, tb, M, N, K)
        pad = tuple(int((-x) % 4 + 4 * rng.integers(0, 3)) for x in (ac, bc, cc))
    else:                                         # arbitrary padding: mostly TMA-ineligible
        pad = tuple(int(x) for x in rng.integers(0, 8, size=3))
    return maj, ta, tb, M, N, K, alpha, beta, pad


@pytest.mark.parametrize("chunk", range(4))
def test_auto_dispatch_fuzz(u, chunk):
    """mode=auto through the drop-in host-pointer entry point; the sweep must exercise all three rule branches."""
    rng = np.random.default_rng(1000 + chunk)
    seen = {"3xtf32": 0, "3xtf32+repack": 0, "simt": 0}
    worst = 0.0
    for i in range(60):
        maj, ta, tb, M, N, K, alpha, beta, pad = _draw(rng, i)
        lo, hi = ((0.0, 1.0), (-0.5, 0.5))[i % 2]
        e = check_case(u, "auto", maj, ta, tb, M, N, K, alpha, beta, pad, seed=7 * chunk + i + 1, lo=lo, hi=hi,
                       naive=(M * N * K <= 1 << 21))
        worst = max(worst, e)
        if alpha != 0.0:
            k = u.last_kernel()
            seen[k + "+repack" if u.last_repacked() else k] += 1
    print(f"fuzz chunk {chunk}: worst relerr {worst:.3e}, kernels {seen}")
    assert seen["3xtf32"] and seen["simt"], seen


@pytest.mark.parametrize("mode", ["simt", "3xtf32"])
def test_forced_modes_fuzz(u, mode):
    """Forced modes: K2 takes anything; K1 takes every TMA-eligible layout down to a single ragged tile."""
    rng = np.random.default_rng(77 if mode == "simt" else 78)
    worst = 0.0
    for i in range(48):
        maj, ta, tb, M, N, K, alpha, beta, pad = _draw(rng, 3 * i + 1 if mode == "3xtf32" else i)
        if mode == "3xtf32" and alpha == 0.0:
            alpha = 1.25
        e = check_case(u, mode, maj, ta, tb, M, N, K, alpha, beta, pad, seed=100 + i, naive=(M * N * K <= 1 << 21))
        worst = max(worst, e)
        if alpha != 0.0 and K > 0:
            assert u.last_kernel() == mode
    print(f"forced {mode} fuzz: worst relerr {worst:.3e}")


def _dev_case(u, mode, ta, tb, M, N, K, alpha, beta, lda, ldb, ldc, oa, ob, oc, seed):
    """Row-major problem on device buffers whose operand base pointers are shifted by oa/ob/oc floats."""
    (ar, ac), (br, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
    A = O.fill_uniform(ar * lda, seed, -0.5, 0.5)
    B = O.fill_uniform(br * ldb, seed + 1, -0.5, 0.5)
    Cm = O.fill_uniform(M * ldc, seed + 2, 0.0, 1.0)
    dA, dB, dC = u.DeviceBuffer(A.size + oa + 4), u.DeviceBuffer(B.size + ob + 4), u.DeviceBuffer(Cm.size + oc + 4)
    try:
        guard = np.full(4, -9.0, np.float32)
        dA.upload(np.concatenate([np.zeros(oa, np.float32), A]))
        dB.upload(np.concatenate([np.zeros(ob, np.float32), B]))
        dC.upload(np.concatenate([np.full(oc, -9.0, np.float32), Cm, guard]))
        u.sgemm_cuda_dev(mode, None, "R", ta, tb, M, N, K, alpha, dA.ptr + 4 * oa, lda, dB.ptr + 4 * ob, ldb, beta,
                         dC.ptr + 4 * oc, ldc)
        u.sync()
        raw = dC.download(oc + Cm.size + 4)
    finally:
        dA.free(); dB.free(); dC.free()
    assert np.all(raw[:oc] == -9.0) and np.all(raw[oc + Cm.size:] == -9.0), "wrote outside C"
    got = np.ascontiguousarray(raw[oc:oc + Cm.size])
    want = O.run14(O.oracle().oracle_sgemm_banded, "R", ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=4)
    e = O.relerr("R", M, N, want, got, ldc)
    assert e <= TOL, f"{mode} {ta}{tb} {M}x{N}x{K} offsets {(oa, ob, oc)} ld {(lda, ldb, ldc)}: relerr {e:.3e}"
    if ldc > N:
        assert np.array_equal(got.reshape(M, ldc)[:, N:], Cm.reshape(M, ldc)[:, N:]), "ld padding of C was written"
    return e


@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
def test_device_pointers_only_4_byte_aligned(u, ta, tb):
    """sgemm_cuda_dev on sub-views: A/B bases off the 16-byte grid make the problem TMA-ineligible (auto: repack when
    large, K2 otherwise; forced K1: an error, never a fallback); a misaligned C must switch both kernels' epilogues to
    scalar stores and still never touch a byte outside the M x N region."""
    M, N, K = 300, 260, 100
    (_, ac), (_, bc), _ = O.stored_shapes("R", ta, tb, M, N, K)
    for (oa, ob, oc) in ((1, 0, 0), (0, 3, 0), (0, 0, 1), (2, 1, 3)):
        _dev_case(u, "auto", ta, tb, M, N, K, 1.5, 0.5, ac + 4, bc + 4, N + 4, oa, ob, oc, seed=31)
        if oa or ob:
            assert u.last_kernel() == "3xtf32" and u.last_repacked()
        _dev_case(u, "simt", ta, tb, M, N, K, 1.5, 0.5, ac + 4, bc + 4, N + 4, oa, ob, oc, seed=32)
        # small problem: rule (3), K2 on the misaligned views directly
        _dev_case(u, "auto", ta, tb, 70, 90, 40, 1.0, 0.0, (70 if ta == "T" else 40) + 1, (40 if tb == "T" else 90) + 2, 93, oa, ob, oc, seed=33)
        assert u.last_kernel() == "simt"
    # K1 with aligned operands and a misaligned C: scalar epilogue
    _dev_case(u, "3xtf32", ta, tb, M, N, K, 1.5, 0.5, ac + 4, bc + 4, N + 3, 0, 0, 1, seed=34)
    assert u.last_kernel() == "3xtf32"
    # forced K1 on a misaligned operand is an error and leaves C untouched
    with pytest.raises(u.UgemmCudaError):
        _dev_case(u, "3xtf32", ta, tb, M, N, K, 1.0, 0.0, ac + 4, bc + 4, N + 4, 1, 0, 0, seed=35)
    u.clear_error() if hasattr(u, "clear_error") else None
