"""compute-sanitizer cases for the beta != 0 path of the TS kernel: the old C arrives through the TMA unit in each epilogue warp's
staging box (the box the tile stores leave from) and is added between promotions.  Several tiles per CTA (the box alternates between
C loads and tile stores), ragged edges (clipped boxes), both CTA-group modes, the stream-K tail beside it, strided batch, and the
register path (C the TMA unit cannot address: ldc % 4 != 0)."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import ugemm_b200 as u
rng = np.random.default_rng(1)
def case(ta, tb, M, N, K, padc=0, alpha=1.5, beta=0.5, cg=0, mode="3xtf32"):
    u.set_k1_tuning(cta_group=cg)
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    A = rng.uniform(0, 1, (ar, ac)).astype(np.float32); B = rng.uniform(0, 1, (br, bc)).astype(np.float32)
    Cm = rng.uniform(-300, 300, (M, N + padc)).astype(np.float32); C0 = Cm.copy()
    fn = {"auto": u.sgemm_cuda, "3xtf32": u.sgemm_cuda_3xtf32}[mode]
    fn("R", ta, tb, M, N, K, alpha, A.ravel(), ac, B.ravel(), bc, beta, Cm.ravel(), N + padc)
    opA = A if ta == "N" else A.T; opB = B if tb == "N" else B.T
    ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0[:, :N]
    e = np.linalg.norm(Cm[:, :N] - ref) / np.linalg.norm(ref)
    print(ta, tb, M, N, K, "padc", padc, "cg", cg, "beta", beta, "kernel", u.last_kernel(), "relerr %.2e" % e, flush=True)
    assert e < 1e-5 and np.array_equal(Cm[:, N:], C0[:, N:])
u.set_sm_limit(4)
case("N", "N", 768, 768, 512, cg=2)             # 9 pair tiles on 2 pairs
case("T", "T", 640, 384, 320, padc=4, cg=1)     # 15 single-CTA tiles on 4 CTAs
case("N", "T", 700, 500, 96, padc=8, cg=2, beta=1.0)   # short K: hand-overs run out before the four boxes are in
u.set_sm_limit(0)
case("N", "N", 1100, 900, 1024, cg=2)           # stream-K tail beside whole tiles
case("T", "N", 300, 260, 100, padc=4, cg=2, beta=-2.0)
case("N", "N", 300, 260, 100, padc=3, cg=2)     # ldc % 4 != 0: the register path and the strided stores
u.sgemm_cuda_finish()
print("ok")
