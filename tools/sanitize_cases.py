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
# TS kernel: several tiles per CTA pair (the slice-buffer ring wraps, the pipelined epilogue re-arms groups and takes early
# hand-overs), the stream-K tail with TMA-stored parts + fix-up pass, single-CTA tiles, and the round-1 SS kernel beside it
u.set_sm_limit(4)
case("3xtf32", "N", "N", 768, 768, 512, (0, 0, 0), cg=2)
case("3xtf32", "T", "T", 640, 384, 320, (0, 0, 4), beta=0.0, cg=1)
u.set_sm_limit(0)
case("3xtf32", "N", "N", 1100, 900, 1024, (0, 0, 0), cg=2)
u.set_k1_variant(1)
u.set_sm_limit(4)
case("3xtf32", "N", "N", 768, 768, 512, (0, 0, 0), cg=2)      # the round-1 kernel with its tile-index slots reused, too
u.set_sm_limit(0)
case("3xtf32", "N", "T", 300, 260, 100, (0, 0, 0), cg=2)
case("3xtf32", "T", "N", 132, 260, 36, (4, 0, 4), cg=1)
u.set_k1_variant(0)
case("3xtf32", "N", "N", 512, 512, 256, (0, 0, 0), beta=0.0, cg=2)
x = rng.uniform(-1, 1, 8 * 14 * 14).astype(np.float32); w = rng.uniform(-1, 1, 16 * 8 * 9).astype(np.float32); b = rng.uniform(-1, 1, 16).astype(np.float32)
out = np.zeros(16 * 14 * 14, np.float32)
u.convolution_cuda_LReLU(x, 8, 14, 14, w, 3, 1, 1, out, 16, b)
print("conv ok", float(np.abs(out).sum()) > 0)
# K2's interior fast path (aligned, whole tiles) and the narrow / 256x64 tiles
case("simt", "N", "N", 256, 256, 64, (0, 0, 0))
case("simt", "T", "T", 256, 128, 48, (0, 0, 0))
case("simt", "N", "N", 300, 16, 70, (0, 0, 0))
case("simt", "N", "N", 512, 24, 40, (3, 1, 5))
# level 1 / level 2
for n, ix, iy in ((1000, 1, 1), (1027, 1, 1), (333, 2, 3)):
    xs = rng.uniform(-1, 1, (n - 1) * ix + 1).astype(np.float32); ys = rng.uniform(-1, 1, (n - 1) * iy + 1).astype(np.float32)
    want = ys.copy(); want[::iy] = np.float32(0.5) * xs[::ix] + want[::iy]
    u.saxpy_cuda(n, 0.5, xs, ix, ys, iy)
    assert np.abs(ys - want).max() < 1e-6
for trans, M, N, pad in (("N", 300, 200, 0), ("N", 512, 4100, 0), ("N", 130, 77, 3), ("T", 300, 200, 0), ("T", 3, 5000, 4), ("T", 17, 33, 1)):
    lines, cols = (N, M) if trans == "N" else (M, N)
    Am = rng.uniform(0, 1, (lines, cols + pad)).astype(np.float32); xv = rng.uniform(-1, 1, N).astype(np.float32); yv = rng.uniform(0, 1, M).astype(np.float32)
    op = Am[:, :cols].T if trans == "N" else Am[:, :cols]
    want = 1.5 * (op.astype(np.float64) @ xv) + 0.5 * yv
    u.sgemv_cuda(trans, M, N, 1.5, Am.ravel(), cols + pad, xv, 1, 0.5, yv, 1)
    e = np.linalg.norm(yv - want) / np.linalg.norm(want)
    print("sgemv", trans, M, N, pad, "relerr %.2e" % e, flush=True)
    assert e < 1e-5
# implicit-GEMM convolution (4-D TMA gather), stride 1 and 2, two images; DGEMM
for (ich, h, w, k, pad, ch, nimg, stride) in ((40, 20, 36, 3, 1, 130, 2, 1), (32, 17, 33, 3, 1, 64, 1, 2)):
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
            ref += np.einsum("oc,nchw->nohw", W[:, :, ki, kj], xp[:, :, ki:ki + stride * ho:stride, kj:kj + stride * wo:stride])
    e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print("conv fused", u.last_conv_fused(), (ich, h, w, k, pad, ch, nimg, stride), "relerr %.2e" % e, flush=True)
    assert u.last_conv_fused() and e < 1e-5
u.set_conv_fusion(-1)
for ta, tb, M, N, K in (("N", "N", 130, 70, 33), ("T", "T", 128, 64, 16), ("N", "T", 17, 9, 5)):
    ar, ac = (M, K) if ta == "N" else (K, M); br, bc = (K, N) if tb == "N" else (N, K)
    Ad = rng.uniform(0, 1, (ar, ac)); Bd = rng.uniform(0, 1, (br, bc)); Cd = rng.uniform(0, 1, (M, N)); C0 = Cd.copy()
    u.dgemm_cuda("R", ta, tb, M, N, K, 1.5, Ad.ravel(), ac, Bd.ravel(), bc, 0.5, Cd.ravel(), N)
    ref = 1.5 * ((Ad if ta == "N" else Ad.T) @ (Bd if tb == "N" else Bd.T)) + 0.5 * C0
    e = np.linalg.norm(Cd - ref) / np.linalg.norm(ref)
    print("dgemm", ta, tb, M, N, K, "relerr %.2e" % e, flush=True)
    assert e < 2e-14
u.sgemm_cuda_finish()
