"""compute-sanitizer cases for the code added last: K1's TMA-store epilogue (GEMM 3-D map, fused convolution 4-D map),
the older strided-store path beside it (misaligned C), and sgemm_cuda_mgpu on a 1 x 1 grid."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import ugemm_b200 as u
rng = np.random.default_rng(0)
def case(mode, ta, tb, M, N, K, pad=(0, 0, 0), alpha=1.5, beta=0.5, cg=0, mg=False):
    u.set_k1_tuning(cta_group=cg)
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    A = rng.uniform(0, 1, (ar, ac + pad[0])).astype(np.float32); B = rng.uniform(0, 1, (br, bc + pad[1])).astype(np.float32)
    Cm = rng.uniform(0, 1, (M, N + pad[2])).astype(np.float32); C0 = Cm.copy()
    if mg:
        u.sgemm_cuda_mgpu("R", ta, tb, M, N, K, alpha, A.ravel(), A.shape[1], B.ravel(), B.shape[1], beta, Cm.ravel(), Cm.shape[1], 1, 1, 2)
    else:
        fn = {"auto": u.sgemm_cuda, "3xtf32": u.sgemm_cuda_3xtf32, "simt": u.sgemm_cuda_simt}[mode]
        fn("R", ta, tb, M, N, K, alpha, A.ravel(), A.shape[1], B.ravel(), B.shape[1], beta, Cm.ravel(), Cm.shape[1])
    opA = A[:, :ac] if ta == "N" else A[:, :ac].T; opB = B[:, :bc] if tb == "N" else B[:, :bc].T
    ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0[:, :N]
    e = np.linalg.norm(Cm[:, :N] - ref) / np.linalg.norm(ref)
    print("mgpu" if mg else mode, ta, tb, M, N, K, pad, "cg", cg, "kernel", u.last_kernel(), "relerr %.2e" % e, flush=True)
    assert e < 1e-5 and np.array_equal(Cm[:, N:], C0[:, N:])
case("3xtf32", "N", "N", 300, 260, 100, (0, 0, 0), cg=2)                 # TMA store, ragged boxes, beta preloaded
case("3xtf32", "T", "N", 132, 260, 36, (4, 0, 4), cg=1)                  # TMA store through a padded ldc
case("3xtf32", "N", "T", 512, 512, 256, (0, 0, 0), beta=0.0, cg=2)       # whole tiles
case("3xtf32", "N", "N", 300, 261, 100, (0, 3, 2), cg=2)                 # ldc = 263: strided-store path beside it
u.sgemm_cuda_mgpu_init(1)
case("auto", "N", "N", 300, 260, 100, (0, 0, 4), mg=True)
case("auto", "T", "T", 70, 33, 20, (1, 2, 3), mg=True)
u.sgemm_cuda_mgpu_finish()
for (ich, h, w, k, pad, ch, nimg, stride) in ((40, 20, 36, 3, 1, 130, 2, 1),):
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    xi = rng.uniform(-.5, .5, nimg * ich * h * w).astype(np.float32); wg = rng.uniform(-.5, .5, ch * ich * k * k).astype(np.float32)
    dx, dw, do = u.DeviceBuffer(xi.size).upload(xi), u.DeviceBuffer(wg.size).upload(wg), u.DeviceBuffer(nimg * ch * ho * wo)
    u.set_conv_fusion(1)
    u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, stride, do, ch, None, 1.0, None)
    u.sync()
    got = do.download().reshape(nimg, ch, ho, wo).astype(np.float64)
    xp = np.zeros((nimg, ich, h + 2 * pad, w + 2 * pad)); xp[:, :, pad:pad + h, pad:pad + w] = xi.reshape(nimg, ich, h, w)
    W = wg.reshape(ch, ich, k, k).astype(np.float64); ref = np.zeros((nimg, ch, ho, wo))
    for ki in range(k):
        for kj in range(k):
            ref += np.einsum("oc,ncyx->noyx", W[:, :, ki, kj], xp[:, :, ki:ki + ho, kj:kj + wo])
    e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print("conv fused", u.last_conv_fused(), "relerr %.2e" % e, flush=True)
    assert e < 1e-5
print("all ok")
