"""Small K1 / K2 / conv cases for compute-sanitizer (memcheck, racecheck, synccheck, initcheck)."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import ugemm_b200 as u
rng = np.random.default_rng(0)
def case(mode, ta, tb, M, N, K, pad=(0, 0, 0), alpha=1.5, beta=0.5, cg=0):
    u.set_k1_tuning(cta_group=cg)
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    A = rng.uniform(0, 1, (ar, ac + pad[0])).astype(np.float32); B = rng.uniform(0, 1, (br, bc + pad[1])).astype(np.float32)
    Cm = rng.uniform(0, 1, (M, N + pad[2])).astype(np.float32); C0 = Cm.copy()
    fn = {"auto": u.sgemm_cuda, "3xtf32": u.sgemm_cuda_3xtf32, "simt": u.sgemm_cuda_simt}[mode]
    fn("R", ta, tb, M, N, K, alpha, A.ravel(), A.shape[1], B.ravel(), B.shape[1], beta, Cm.ravel(), Cm.shape[1])
    opA = A[:, :ac] if ta == "N" else A[:, :ac].T; opB = B[:, :bc] if tb == "N" else B[:, :bc].T
    ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0[:, :N]
    e = np.linalg.norm(Cm[:, :N] - ref) / np.linalg.norm(ref)
    print(mode, ta, tb, M, N, K, pad, "cg", cg, "kernel", u.last_kernel(), "relerr %.2e" % e, flush=True)
    assert e < 1e-5
for ta in "NT":
    for tb in "NT":
        case("3xtf32", ta, tb, 300, 260, 100, (0, 0, 0), cg=2)
        case("3xtf32", ta, tb, 132, 260, 36, (4, 0, 4), cg=1)
        case("simt", ta, tb, 129, 97, 131, (3, 5, 7))
case("auto", "N", "N", 300, 257, 100, (1, 2, 3))        # repack path
case("3xtf32", "N", "N", 512, 512, 256, (0, 0, 0), beta=0.0, cg=2)
x = rng.uniform(-1, 1, 8 * 14 * 14).astype(np.float32); w = rng.uniform(-1, 1, 16 * 8 * 9).astype(np.float32); b = rng.uniform(-1, 1, 16).astype(np.float32)
out = np.zeros(16 * 14 * 14, np.float32)
u.convolution_cuda_LReLU(x, 8, 14, 14, w, 3, 1, 1, out, 16, b)
print("conv ok", float(np.abs(out).sum()) > 0)
u.sgemm_cuda_finish()
