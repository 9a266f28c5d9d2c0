"""ctypes loaders for the CHECKER libraries (oracle/liboracle.so, oracle/_ref/libugemm_ref.so).

Test infrastructure only: nothing under ugemm_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libugemm_ref.so")
REF_CONV_SO = os.path.join(ORACLE_DIR, "_ref", "libugemm_ref_conv.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_SIG14 = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_float,
          _f32p, C.c_int, _f32p, C.c_int, C.c_float, _f32p, C.c_int]


_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_DSIG14 = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double,
           _f64p, C.c_int, _f64p, C.c_int, C.c_double, _f64p, C.c_int]
_AXPY = [C.c_int, C.c_float, _f32p, C.c_int, _f32p, C.c_int]
_GEMV = [C.c_char, C.c_int, C.c_int, C.c_float, _f32p, C.c_int, _f32p, C.c_int, C.c_float, _f32p, C.c_int]


def build_oracle(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(ORACLE_DIR, "sgemm_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(src) > os.path.getmtime(ORACLE_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    def stale(so, *srcs):
        return not os.path.exists(so) or any(os.path.getmtime(os.path.join(ORACLE_DIR, s)) > os.path.getmtime(so) for s in srcs)
    if os.path.exists(os.environ.get("UGEMM_REF", "/root/reference") + "/ugemm.h") and (
            force or stale(REF_SO, "ref_shim.c") or stale(REF_CONV_SO, "ref_conv_shim.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None
_ref_conv = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_sgemm_naive.argtypes = _SIG14
        lib.oracle_sgemm_naive.restype = None
        lib.oracle_sgemm_banded.argtypes = [C.c_int] + _SIG14
        lib.oracle_sgemm_banded.restype = None
        lib.oracle_relerr.argtypes = [C.c_char, C.c_int, C.c_int, _f32p, _f32p, C.c_int]
        lib.oracle_relerr.restype = C.c_double
        lib.oracle_cmp_results.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_int,
                                           np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]
        lib.oracle_cmp_results.restype = C.c_int
        lib.oracle_fill_uniform.argtypes = [_f32p, C.c_size_t, C.c_uint64, C.c_float, C.c_float]
        lib.oracle_fill_uniform.restype = None
        lib.oracle_fill_uniform_at.argtypes = [_f32p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
        lib.oracle_fill_uniform_at.restype = None
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_im2col.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        lib.oracle_im2col.restype = None
        lib.oracle_convolution.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p,
                                           C.c_int, C.c_void_p, C.c_float, _f32p]
        lib.oracle_convolution.restype = None
        lib.oracle_dgemm_naive.argtypes = _DSIG14
        lib.oracle_dgemm_naive.restype = None
        lib.oracle_relerr_f64.argtypes = [C.c_int, C.c_int, _f64p, _f64p, C.c_int]
        lib.oracle_relerr_f64.restype = C.c_double
        lib.oracle_saxpy.argtypes = _AXPY
        lib.oracle_saxpy.restype = None
        lib.oracle_sgemv.argtypes = _GEMV
        lib.oracle_sgemv.restype = None
        _oracle = lib
    return _oracle


def ref():
    """The UNMODIFIED reference compiled into oracle/_ref (None if it was never built)."""
    global _ref
    if _ref is None:
        build_oracle()
        if not os.path.exists(REF_SO):
            return None
        lib = C.CDLL(REF_SO)
        for name in ("ref_sgemm_cpu", "ref_sgemm_c", "ref_sgemm_avx", "ref_sgemm_sse"):
            getattr(lib, name).argtypes = _SIG14
            getattr(lib, name).restype = None
        lib.ref_sgemm_avx_mt.argtypes = [C.c_int] + _SIG14
        lib.ref_sgemm_avx_mt.restype = None
        lib.ref_max_threads.restype = C.c_int
        if hasattr(lib, "ref_saxpy_cpu"):   # a prebuilt _ref from before these exports existed still serves the GEMM tests
            lib.ref_saxpy_cpu.argtypes = _AXPY
            lib.ref_saxpy_cpu.restype = None
            lib.ref_sgemv_cpu.argtypes = _GEMV
            lib.ref_sgemv_cpu.restype = None
        if hasattr(lib, "ref_dgemm_cpu"):
            for name in ("ref_dgemm_cpu", "ref_dgemm_c", "ref_dgemm_avx"):
                getattr(lib, name).argtypes = _DSIG14
                getattr(lib, name).restype = None
        _ref = lib
    return _ref


def ref_conv():
    """The reference's UNMODIFIED convolution host code (sgemm_gl1.h im2col), compiled against GL stubs (None if never built)."""
    global _ref_conv
    if _ref_conv is None:
        build_oracle()
        if not os.path.exists(REF_CONV_SO):
            return None
        lib = C.CDLL(REF_CONV_SO)
        lib.ref_im2col.argtypes = [_f32p] + [C.c_int] * 9 + [_f32p]
        lib.ref_im2col.restype = None
        lib.ref_gl_convolution.argtypes = [C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_void_p]
        lib.ref_gl_convolution.restype = C.c_int
        _ref_conv = lib
    return _ref_conv


def b(ch):
    return ch.encode() if isinstance(ch, str) else ch


def fill_uniform(n, seed, lo=0.0, hi=1.0):
    x = np.empty(int(n), dtype=np.float32)
    oracle().oracle_fill_uniform(x, x.size, seed, lo, hi)
    return x


def fill_uniform_at(n, seed, offset, lo=0.0, hi=1.0):
    """Elements [offset, offset + n) of the stream fill_uniform(., seed) produces (a window of a device-generated matrix)."""
    x = np.empty(int(n), dtype=np.float32)
    oracle().oracle_fill_uniform_at(x, x.size, seed, int(offset), lo, hi)
    return x


def stored_shapes(major, ta, tb, M, N, K):
    """(rows, cols) of A, B, C as stored, i.e. (count of ld-strided lines, line length)."""
    if major == "R":
        a = (M, K) if ta == "N" else (K, M)
        bb = (K, N) if tb == "N" else (N, K)
        c = (M, N)
    else:
        a = (K, M) if ta == "N" else (M, K)
        bb = (N, K) if tb == "N" else (K, N)
        c = (N, M)
    return a, bb, c


def make_problem(major, ta, tb, M, N, K, pad=(0, 0, 0), seed=1, lo=0.0, hi=1.0, sentinel=None):
    """Seeded A, B, C buffers (1-D float32, ld-strided) + their leading dimensions."""
    (ar, ac), (br, bc), (cr, cc) = stored_shapes(major, ta, tb, M, N, K)
    lda, ldb, ldc = ac + pad[0], bc + pad[1], cc + pad[2]
    A = fill_uniform(max(ar * lda, 1), seed * 3 + 0, lo, hi)
    B = fill_uniform(max(br * ldb, 1), seed * 3 + 1, lo, hi)
    Cm = fill_uniform(max(cr * ldc, 1), seed * 3 + 2, lo, hi)
    if sentinel is not None and pad[2]:
        Cm.reshape(cr, ldc)[:, cc:] = sentinel
    return A, lda, B, ldb, Cm, ldc


def relerr(major, M, N, ref_c, res_c, ld):
    return float(oracle().oracle_relerr(b(major), M, N, ref_c, res_c, ld))


def run14(fn, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=None):
    out = Cm.copy()
    args = (b(major), b(ta), b(tb), M, N, K, alpha, A, lda, B, ldb, beta, out, ldc)
    if threads is None:
        fn(*args)
    else:
        fn(threads, *args)
    return out


def make_problem_f64(major, ta, tb, M, N, K, pad=(0, 0, 0), seed=1, lo=0.0, hi=1.0, sentinel=None):
    """float64 twin of make_problem: the seeded fp32 stream widened to double plus a second stream scaled by 2^-24, so
    the low mantissa bits are populated too (a DGEMM that secretly computed in fp32 would fail the gate)."""
    (ar, ac), (br, bc), (cr, cc) = stored_shapes(major, ta, tb, M, N, K)
    lda, ldb, ldc = ac + pad[0], bc + pad[1], cc + pad[2]

    def stream(n, s):
        return fill_uniform(max(n, 1), s, lo, hi).astype(np.float64) + fill_uniform(max(n, 1), s + 7919, 0.0, 1.0).astype(np.float64) * 2.0 ** -24

    A, B, Cm = stream(ar * lda, seed * 3), stream(br * ldb, seed * 3 + 1), stream(cr * ldc, seed * 3 + 2)
    if sentinel is not None and pad[2]:
        Cm.reshape(cr, ldc)[:, cc:] = sentinel
    return A, lda, B, ldb, Cm, ldc
