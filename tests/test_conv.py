"""Convolution callers of the GEMM: im2col and convolution(+bias+LeakyReLU).

CPU part: the oracle's restatement of the reference's im2col / gl_convolution_LReLU (sgemm_gl1.h:166-218) is pinned
against a definition-level direct convolution (the headers that hold the reference versions need GLFW/OpenCL and do
not compile here).  GPU part (-m gpu): the CUDA entry points against that oracle, bit-exact for im2col (pure data
movement), normwise relerr <= 1e-5 for the convolutions."""
import numpy as np
import pytest

import _oracle as O

GEOMS = [  # ich, h, w, k, pad, stride, ch
    (1, 5, 5, 3, 1, 1, 2),
    (3, 8, 10, 3, 0, 1, 4),
    (4, 9, 7, 2, 1, 2, 5),
    (8, 14, 14, 3, 1, 1, 16),
    (2, 6, 6, 5, 2, 1, 3),
    (3, 12, 9, 4, 2, 2, 6),
]


def direct_conv(x, wgt, k, pad, stride):
    """outputs[co, io, jo] = sum_{c,ki,kj} wgt[co, c, ki, kj] * x_padded[c, io*stride+ki, jo*stride+kj]   (fp64)"""
    ich, h, w = x.shape
    ch = wgt.shape[0]
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    xp = np.zeros((ich, h + 2 * pad, w + 2 * pad))
    xp[:, pad:pad + h, pad:pad + w] = x
    out = np.zeros((ch, ho, wo))
    for ki in range(k):
        for kj in range(k):
            patch = xp[:, ki:ki + stride * ho:stride, kj:kj + stride * wo:stride]      # ich, ho, wo
            out += np.einsum("oc,chw->ohw", wgt[:, :, ki, kj].astype(np.float64), patch)
    return out


def make_conv(ich, h, w, k, ch, seed):
    x = O.fill_uniform(ich * h * w, seed, -0.5, 0.5)
    wgt = O.fill_uniform(ch * ich * k * k, seed + 1, -0.5, 0.5)
    bias = O.fill_uniform(ch, seed + 2, -0.5, 0.5)
    return x, wgt, bias


def oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, bias, slope):
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = np.zeros(ch * ho * wo, np.float32)
    ws = np.zeros(ich * k * k * ho * wo, np.float32)
    O.oracle().oracle_convolution(2, x, ich, w, h, wgt, k, pad, stride, out, ch,
                                  None if bias is None else bias.ctypes.data, slope, ws)
    return out, ws


@pytest.mark.parametrize("geom", GEOMS)
def test_oracle_im2col_and_convolution_match_the_definition(geom):
    ich, h, w, k, pad, stride, ch = geom
    x, wgt, bias = make_conv(ich, h, w, k, ch, seed=50)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out, col = oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, None, 1.0)
    # im2col: every entry is either an image pixel or a padding zero, at the documented position
    xp = np.zeros((ich, h + 2 * pad, w + 2 * pad), np.float32)
    xp[:, pad:pad + h, pad:pad + w] = x.reshape(ich, h, w)
    colm = col.reshape(ich, k, k, ho, wo)
    for ki in range(k):
        for kj in range(k):
            assert np.array_equal(colm[:, ki, kj], xp[:, ki:ki + stride * ho:stride, kj:kj + stride * wo:stride])
    ref = direct_conv(x.reshape(ich, h, w), wgt.reshape(ch, ich, k, k), k, pad, stride).reshape(-1)
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) <= 1e-6
    # + bias + LeakyReLU(0.1), sgemm_gl1.h:210-217
    out2, _ = oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, bias, 0.1)
    r2 = ref.reshape(ch, -1) + bias[:, None].astype(np.float64)
    r2 = np.where(r2 > 0, r2, 0.1 * r2).reshape(-1)
    assert np.linalg.norm(out2 - r2) / np.linalg.norm(r2) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOMS + [(128, 56, 56, 3, 1, 1, 256), (64, 57, 41, 3, 1, 2, 96)])
def test_gpu_im2col_and_convolution(u, geom):
    """(128,56,56,3,1,1,256) is one image of BASELINE config 4: M=256, N=3136, K=1152 -> K1 (3xTF32)."""
    ich, h, w, k, pad, stride, ch = geom
    x, wgt, bias = make_conv(ich, h, w, k, ch, seed=70)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    want, want_col = oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, None, 1.0)
    col = np.full(ich * k * k * ho * wo, np.nan, np.float32)
    u.im2col_cuda(x, ich, h, w, k, pad, stride, col)
    assert np.array_equal(col, want_col), "im2col is pure data movement: must be bit-exact"
    out = np.full(ch * ho * wo, np.nan, np.float32)
    u.convolution_cuda(x, ich, w, h, wgt, k, pad, stride, out, ch)
    e = np.linalg.norm(out.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
    assert e <= 1e-5, e
    big = ch >= 128 and ho * wo >= 128 and (ho * wo) % 4 == 0 and (ich * k * k) % 4 == 0 and ich * k * k >= 32
    assert u.last_kernel() == ("3xtf32" if big else "simt")
    want2, _ = oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, bias, 0.1)
    out2 = np.full(ch * ho * wo, np.nan, np.float32)
    u.convolution_cuda_LReLU(x, ich, w, h, wgt, k, pad, stride, out2, ch, bias)
    e2 = np.linalg.norm(out2.astype(np.float64) - want2) / np.linalg.norm(want2.astype(np.float64))
    assert e2 <= 1e-5, e2
    # the activation really happened: negatives are scaled by 0.1 relative to the plain result + bias
    plain = out.reshape(ch, -1) + bias[:, None]
    neg = plain < -1e-3
    if neg.any():
        assert np.allclose(out2.reshape(ch, -1)[neg], 0.1 * plain[neg], rtol=1e-3, atol=1e-5)


FUSED_GEOMS = [  # ich, h, w, k, pad, ch, nimg, stride
    (32, 8, 32, 3, 1, 128, 1, 1),      # one 32-pixel chunk per output row, exact channel block
    (64, 14, 64, 3, 1, 256, 1, 1),     # 2-CTA-sized M
    (128, 56, 56, 3, 1, 256, 1, 1),    # one image of BASELINE config 4 (output width 56 -> padded to 64)
    (40, 20, 36, 3, 0, 130, 2, 1),     # ragged channels (40 -> 64), ragged filters, wo = 34, no padding, two images
    (24, 12, 28, 5, 2, 96, 3, 1),      # 5x5, ich < 32
    (3, 16, 32, 3, 1, 64, 2, 1),       # first-layer-like: 3 channels
    (96, 9, 12, 1, 0, 64, 4, 1),       # 1x1 convolution, wo = 12
    (16, 10, 8, 2, 1, 70, 1, 1),       # even kernel, wo = 9
    (64, 15, 30, 3, 1, 96, 2, 1),      # width not a multiple of 4, odd height
    (30, 11, 57, 3, 1, 128, 1, 1),     # odd width, channels not a multiple of 4 (30 -> cs 32)
    (6, 7, 9, 3, 2, 64, 2, 1),         # pad > (k-1)/2: output larger than the input
    (64, 57, 41, 3, 1, 96, 2, 2),      # stride 2 (TMA element stride), odd sizes
    (32, 64, 64, 3, 1, 128, 1, 2),     # stride 2, wo = 32
    (48, 33, 70, 5, 2, 64, 1, 3),      # stride 3, 5x5
    (32, 16, 16, 2, 0, 64, 2, 2),      # 2x2 stride 2 (pooling-like), wo = 8
]


@pytest.mark.gpu
@pytest.mark.parametrize("geom", FUSED_GEOMS)
def test_gpu_implicit_gemm_convolution(u, geom):
    """The fused (implicit-GEMM, 4-D TMA gather) convolution against the oracle's im2col + GEMM, image by image, with and
    without bias + LeakyReLU, and against the unfused CUDA path on the same inputs."""
    ich, h, w, k, pad, ch, nimg, stride = geom
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    x = O.fill_uniform(nimg * ich * h * w, 301, -0.5, 0.5)
    wgt = O.fill_uniform(ch * ich * k * k, 302, -0.5, 0.5)
    bias = O.fill_uniform(ch, 303, -0.5, 0.5)
    dx, dw, db = u.DeviceBuffer(x.size).upload(x), u.DeviceBuffer(wgt.size).upload(wgt), u.DeviceBuffer(ch).upload(bias)
    dout, dws = u.DeviceBuffer(nimg * ch * ho * wo), u.DeviceBuffer(ich * k * k * ho * wo)
    try:
        for d_bias, b_host, slope in ((None, None, 1.0), (db, bias, 0.1)):
            want = np.concatenate([oracle_conv(x[i * ich * h * w:(i + 1) * ich * h * w], ich, w, h, wgt, k, pad, stride, ch, b_host, slope)[0]
                                   for i in range(nimg)])
            u.set_conv_fusion(1)
            dout.upload(np.full(dout.n, np.nan, np.float32))
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, stride, dout, ch, d_bias, slope, None)   # no workspace needed
            u.sync()
            assert u.last_conv_fused() and u.last_kernel() == "3xtf32"
            got = dout.download()
            e = np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64))
            assert np.isfinite(got).all() and e <= 1e-5, (geom, slope, e)
            u.set_conv_fusion(0)
            u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, stride, dout, ch, d_bias, slope, dws)
            u.sync()
            assert not u.last_conv_fused()
            ref = dout.download()
            assert np.linalg.norm(got.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64)) <= 1e-5
    finally:
        u.set_conv_fusion(-1)


@pytest.mark.gpu
def test_conv_fusion_rule_and_fallbacks(u):
    """auto rule: fused only for bounded padding waste (and strides <= 8); everything else takes im2col + GEMM."""
    def run(ich, h, w, k, pad, stride, ch):
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        x, wgt, _ = make_conv(ich, h, w, k, ch, seed=90)
        dx, dw = u.DeviceBuffer(x.size).upload(x), u.DeviceBuffer(wgt.size).upload(wgt)
        dout, dws = u.DeviceBuffer(ch * ho * wo), u.DeviceBuffer(ich * k * k * ho * wo)
        u.convolution_cuda_dev("auto", None, dx, ich, w, h, dw, k, pad, stride, dout, ch, None, 1.0, dws)
        u.sync()
        want, _ = oracle_conv(x, ich, w, h, wgt, k, pad, stride, ch, None, 1.0)
        got = dout.download()
        assert np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want.astype(np.float64)) <= 1e-5
        return u.last_conv_fused()
    assert run(128, 56, 56, 3, 1, 1, 256)          # config 4's layer: 64/56 * 128/128 = 1.14 -> fused
    assert not run(64, 57, 41, 3, 1, 2, 96)        # stride 2: wo = 21 -> 32, too much padded work
    assert run(64, 57, 61, 3, 1, 2, 96)            # stride 2: wo = 31 -> 32
    assert not run(64, 100, 100, 3, 1, 9, 96)      # stride 9: beyond the TMA box limit
    assert run(64, 30, 30, 3, 1, 1, 96)            # any width: 32/30 padded columns
    assert not run(128, 13, 12, 3, 1, 1, 256)      # wo = 12 -> 32: too much padded work
    assert not run(3, 32, 32, 3, 1, 1, 64)         # 3 channels -> 32
    assert not run(64, 28, 28, 3, 1, 1, 32)        # few filters


@pytest.mark.gpu
def test_conv_unfused_path_requires_a_workspace(u):
    """A geometry that takes im2col + GEMM with d_workspace == NULL is an error, not a crash; the output is untouched."""
    ich, h, w, k, pad, stride, ch = 8, 20, 20, 3, 1, 9, 16
    x, wgt, _ = make_conv(ich, h, w, k, ch, seed=95)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    dx, dw = u.DeviceBuffer(x.size).upload(x), u.DeviceBuffer(wgt.size).upload(wgt)
    sentinel = np.full(ch * ho * wo, 7.0, np.float32)
    dout = u.DeviceBuffer(sentinel.size).upload(sentinel)
    with pytest.raises(u.UgemmCudaError):
        u.convolution_cuda_dev("auto", None, dx, ich, w, h, dw, k, pad, stride, dout, ch, None, 1.0, None)
    with pytest.raises(u.UgemmCudaError):
        u.convolution_cuda_batched_dev("auto", None, dx, 1, ich, w, h, dw, k, pad, stride, dout, ch, None, 1.0, None)
    u.sync()
    assert np.array_equal(dout.download(), sentinel)
