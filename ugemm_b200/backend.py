"""Host-side mirror of the reference's SGEMM backend interface, bound to libugemm_cuda.so via ctypes.

Names, argument order and error behaviour follow the reference (`uut(major, transA, transB, M, N, K, alpha,
A, lda, B, ldb, beta, C, ldc)`, check_sgemm.c:96-103; `sgemm_init / sgemm_finish / sgemm_rnn / sgemm_rnt /
sgemm_rtn`, sgemm_test.c:19-33).  This module contains NO arithmetic: every call goes through the C ABI of
include/ugemm_cuda.h into the hand-written sm_100a kernels.  If the shared library is missing it is built
with nvcc; if that is impossible the import fails loudly -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MODE_AUTO, MODE_3XTF32, MODE_SIMT = 0, 1, 2
_MODES = {"auto": MODE_AUTO, "3xtf32": MODE_3XTF32, "simt": MODE_SIMT, 0: 0, 1: 1, 2: 2}

_lib = None


class UgemmCudaError(RuntimeError):
    pass


def _ptr(x, dtype=np.float32):
    """Host numpy array / device pointer int / object with data_ptr() -> c_void_p."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, np.ndarray):
        if x.dtype != dtype or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"host buffers must be C-contiguous {np.dtype(dtype).name} numpy arrays")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def lib():
    """Load (building first if needed) libugemm_cuda.so.  Raises if the CUDA extension is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UGEMM_CUDA_LIB") or _build.LIB   # override: A/B-test another build of the same ABI
    if path == _build.LIB and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # stale-but-present library is still usable on a box without nvcc
            if not os.path.exists(path):
                raise UgemmCudaError(f"libugemm_cuda.so is missing and could not be built: {e}") from e
    L = C.CDLL(path)
    sig14 = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_float,
             C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    L.sgemm_cuda_init.argtypes = [C.c_int, C.c_size_t]
    L.sgemm_cuda_init.restype = C.c_int
    L.sgemm_cuda_finish.restype = None
    for n in ("sgemm_cuda", "sgemm_cuda_3xtf32", "sgemm_cuda_simt"):
        getattr(L, n).argtypes = sig14
        getattr(L, n).restype = None
    L.sgemm_cuda_dev.argtypes = [C.c_int, C.c_void_p] + sig14
    L.sgemm_cuda_dev.restype = C.c_int
    bat = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_longlong,
           C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_void_p, C.c_int, C.c_longlong, C.c_int]
    L.sgemm_cuda_batched.argtypes = bat
    L.sgemm_cuda_batched.restype = None
    L.sgemm_cuda_batched_dev.argtypes = [C.c_int, C.c_void_p] + bat
    L.sgemm_cuda_batched_dev.restype = C.c_int
    L.sgemm_cuda_k1_eligible.argtypes = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.sgemm_cuda_k1_eligible.restype = C.c_int
    L.sgemm_cuda_time_dev.argtypes = [C.c_int, C.c_int, C.c_int] + sig14 + [C.POINTER(C.c_float)] * 3
    L.sgemm_cuda_time_dev.restype = C.c_int
    L.sgemm_cuda_last_error.restype = C.c_char_p
    L.sgemm_cuda_clear_error.restype = None
    L.sgemm_cuda_last_kernel.restype = C.c_int
    L.sgemm_cuda_last_repacked.restype = C.c_int
    L.sgemm_cuda_launch_count.restype = C.c_ulonglong
    L.ugemm_cuda_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.c_char_p, C.c_int]
    L.ugemm_cuda_device_info.restype = C.c_int
    L.sgemm_cuda_set_k1_tuning.argtypes = [C.c_int, C.c_int, C.c_int]
    L.sgemm_cuda_set_k1_tuning.restype = None
    L.sgemm_cuda_k1_plan.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)]
    L.sgemm_cuda_k1_plan.restype = C.c_int
    L.sgemm_cuda_k1_plan_item.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.sgemm_cuda_k1_plan_item.restype = C.c_int
    L.sgemm_cuda_set_k1_variant.argtypes = [C.c_int]
    L.sgemm_cuda_set_k1_variant.restype = None
    L.sgemm_cuda_set_sm_limit.argtypes = [C.c_int]
    L.sgemm_cuda_set_sm_limit.restype = None
    L.ugemm_cuda_malloc.argtypes = [C.c_size_t]
    L.ugemm_cuda_malloc.restype = C.c_void_p
    L.ugemm_cuda_free.argtypes = [C.c_void_p]
    L.ugemm_cuda_malloc_host.argtypes = [C.c_size_t]
    L.ugemm_cuda_malloc_host.restype = C.c_void_p
    L.ugemm_cuda_free_host.argtypes = [C.c_void_p]
    L.ugemm_cuda_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ugemm_cuda_memcpy_h2d.restype = C.c_int
    L.ugemm_cuda_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ugemm_cuda_memcpy_d2h.restype = C.c_int
    L.ugemm_cuda_sync.restype = C.c_int
    L.ugemm_cuda_memcpy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.ugemm_cuda_memcpy_async.restype = C.c_int
    L.ugemm_cuda_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    L.ugemm_cuda_ipc_export.restype = C.c_int
    L.ugemm_cuda_ipc_import.argtypes = [C.c_void_p]
    L.ugemm_cuda_ipc_import.restype = C.c_void_p
    L.ugemm_cuda_ipc_close.argtypes = [C.c_void_p]
    L.ugemm_cuda_ipc_close.restype = C.c_int
    PI, PLL, PF = C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_float)
    L.sgemm_cuda_shard_plan.argtypes = [C.c_int] * 5 + [PI] * 6
    L.sgemm_cuda_shard_owners.argtypes = [C.c_int] * 6 + [PI, PI, PLL, PLL]
    L.sgemm_cuda_shard_unique_id.argtypes = [C.c_char_p]
    L.sgemm_cuda_shard_init.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sgemm_cuda_shard_finish.argtypes = []
    L.sgemm_cuda_shard_finish.restype = None
    L.sgemm_cuda_shard_transport.argtypes = []
    L.sgemm_cuda_shard_generate.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_float, C.c_float]
    L.sgemm_cuda_shard_run.argtypes = [C.c_int, C.c_int, C.c_int, PF]
    L.sgemm_cuda_shard_allreduce.argtypes = [PF, C.c_int]
    L.sgemm_cuda_shard_block.argtypes = [C.POINTER(C.c_void_p), PI, PI, PI, PI]
    L.sgemm_cuda_shard_host_buffers.argtypes = [C.POINTER(C.c_void_p), PLL, C.POINTER(C.c_void_p), PLL]
    L.sgemm_cuda_shard_download_owned.argtypes = []
    L.sgemm_cuda_shard_run_host.argtypes = [C.c_int, C.c_int, PF, PLL, PLL]
    L.sgemm_cuda_shard_copy_floor.argtypes = [C.c_int, PF]
    for f in ("plan", "owners", "unique_id", "init", "transport", "generate", "run", "allreduce", "block", "host_buffers", "download_owned", "run_host", "copy_floor"):
        getattr(L, "sgemm_cuda_shard_" + f).restype = C.c_int
    L.sgemm_cuda_mgpu_init.argtypes = [C.c_int]
    L.sgemm_cuda_mgpu_init.restype = C.c_int
    L.sgemm_cuda_mgpu_finish.argtypes = []
    L.sgemm_cuda_mgpu_finish.restype = None
    L.sgemm_cuda_mgpu_count.argtypes = []
    L.sgemm_cuda_mgpu_count.restype = C.c_int
    L.ugemm_cuda_device_count.argtypes = []
    L.ugemm_cuda_device_count.restype = C.c_int
    L.sgemm_cuda_mgpu.argtypes = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.sgemm_cuda_mgpu.restype = C.c_int
    L.ugemm_fill_uniform_host.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_float, C.c_float]
    L.ugemm_fill_uniform_host.restype = None
    L.ugemm_fill_uniform_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_float, C.c_float, C.c_void_p]
    L.ugemm_fill_uniform_dev.restype = C.c_int
    L.ugemm_fill_uniform_host_2d.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint64,
                                             C.c_uint64, C.c_float, C.c_float]
    L.ugemm_fill_uniform_host_2d.restype = None
    L.ugemm_fill_uniform_dev_2d.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint64,
                                            C.c_uint64, C.c_float, C.c_float, C.c_void_p]
    L.ugemm_fill_uniform_dev_2d.restype = C.c_int
    L.im2col_cuda.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.im2col_cuda.restype = None
    L.im2col_cuda_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.im2col_cuda_dev.restype = C.c_int
    conv10 = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.convolution_cuda.argtypes = conv10
    L.convolution_cuda.restype = None
    L.convolution_cuda_LReLU.argtypes = conv10 + [C.c_void_p]
    L.convolution_cuda_LReLU.restype = None
    L.convolution_cuda_dev.argtypes = [C.c_int, C.c_void_p] + conv10 + [C.c_void_p, C.c_float, C.c_void_p]
    L.convolution_cuda_dev.restype = C.c_int
    L.saxpy_cuda.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.saxpy_cuda.restype = None
    L.saxpy_cuda_dev.argtypes = [C.c_void_p] + L.saxpy_cuda.argtypes
    L.saxpy_cuda_dev.restype = C.c_int
    L.sgemv_cuda.argtypes = [C.c_char, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int]
    L.sgemv_cuda.restype = None
    L.sgemv_cuda_dev.argtypes = [C.c_void_p] + L.sgemv_cuda.argtypes
    L.sgemv_cuda_dev.restype = C.c_int
    dsig14 = [C.c_char, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double,
              C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    L.dgemm_cuda.argtypes = dsig14
    L.dgemm_cuda.restype = None
    L.dgemm_cuda_dev.argtypes = [C.c_void_p] + dsig14
    L.dgemm_cuda_dev.restype = C.c_int
    L.dgemm_cuda_time_dev.argtypes = [C.c_int, C.c_int] + dsig14 + [C.POINTER(C.c_float)] * 2
    L.dgemm_cuda_time_dev.restype = C.c_int
    L.convolution_cuda_batched_dev.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
    L.convolution_cuda_batched_dev.restype = C.c_int
    L.sgemm_cuda_set_conv_fusion.argtypes = [C.c_int]
    L.sgemm_cuda_set_conv_fusion.restype = None
    L.sgemm_cuda_last_conv_fused.restype = C.c_int
    L.ugemm_cuda_probe_tf32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ugemm_cuda_probe_tf32.restype = C.c_int
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "sgemm_cuda_init", "sgemm_cuda_finish", "sgemm_cuda", "sgemm_cuda_3xtf32", "sgemm_cuda_simt",
    "sgemm_cuda_dev", "sgemm_cuda_batched", "sgemm_cuda_batched_dev", "sgemm_cuda_k1_eligible", "sgemm_cuda_time_dev", "sgemm_cuda_last_error",
    "sgemm_cuda_clear_error", "sgemm_cuda_last_kernel", "sgemm_cuda_last_repacked", "sgemm_cuda_launch_count", "ugemm_cuda_device_info",
    "sgemm_cuda_set_k1_tuning", "sgemm_cuda_set_k1_variant", "sgemm_cuda_k1_plan", "sgemm_cuda_k1_plan_item", "sgemm_cuda_set_sm_limit", "ugemm_cuda_malloc", "ugemm_cuda_free", "ugemm_cuda_malloc_host",
    "ugemm_cuda_free_host", "ugemm_cuda_memcpy_h2d", "ugemm_cuda_memcpy_d2h", "ugemm_cuda_sync",
    "ugemm_cuda_memcpy_async", "ugemm_cuda_ipc_export", "ugemm_cuda_ipc_import", "ugemm_cuda_ipc_close",
    "sgemm_cuda_mgpu_init", "sgemm_cuda_mgpu_finish", "sgemm_cuda_mgpu_count", "sgemm_cuda_mgpu", "sgemm_cuda_mgpu_run", "sgemm_cuda_mgpu_plan", "ugemm_cuda_device_count",
    "sgemm_cuda_shard_plan", "sgemm_cuda_shard_owners", "sgemm_cuda_shard_unique_id", "sgemm_cuda_shard_init", "sgemm_cuda_shard_finish",
    "sgemm_cuda_shard_transport", "sgemm_cuda_shard_generate", "sgemm_cuda_shard_run", "sgemm_cuda_shard_allreduce", "sgemm_cuda_shard_block",
    "sgemm_cuda_shard_host_buffers", "sgemm_cuda_shard_download_owned", "sgemm_cuda_shard_run_host", "sgemm_cuda_shard_copy_floor",
    "ugemm_fill_uniform_host", "ugemm_fill_uniform_dev", "ugemm_fill_uniform_host_2d", "ugemm_fill_uniform_dev_2d",
    "ugemm_cuda_probe_tf32", "im2col_cuda", "im2col_cuda_dev", "convolution_cuda", "convolution_cuda_LReLU",
    "convolution_cuda_dev", "convolution_cuda_batched_dev", "sgemm_cuda_set_conv_fusion", "sgemm_cuda_last_conv_fused", "saxpy_cuda", "saxpy_cuda_dev", "sgemv_cuda", "sgemv_cuda_dev",
    "dgemm_cuda", "dgemm_cuda_dev", "dgemm_cuda_time_dev",
]


def _b(ch):
    return ch.encode() if isinstance(ch, str) else ch


def last_error():
    e = lib().sgemm_cuda_last_error()
    return e.decode() if e else None


def check():
    """Raise (and clear) the sticky error of the void entry points, if any."""
    e = last_error()
    if e:
        lib().sgemm_cuda_clear_error()
        raise UgemmCudaError(e)


def sgemm_cuda_init(device=-1, arena_bytes=0):
    if lib().sgemm_cuda_init(device, arena_bytes):
        check()


def sgemm_cuda_finish():
    lib().sgemm_cuda_finish()


def _host14(fn, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    fn(_b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(A), lda, _ptr(B), ldb, beta, _ptr(Cm), ldc)
    check()


def sgemm_cuda(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    """Drop-in `uut`: host buffers, blocking, C updated in place (rule-based K1/K2 choice)."""
    _host14(lib().sgemm_cuda, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)


def sgemm_cuda_3xtf32(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    _host14(lib().sgemm_cuda_3xtf32, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)


def sgemm_cuda_simt(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    _host14(lib().sgemm_cuda_simt, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc)


def sgemm_cuda_dev(mode, stream, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc):
    """Device pointers (ints, or anything with .data_ptr()); asynchronous on `stream` (int handle or None)."""
    rc = lib().sgemm_cuda_dev(_MODES[mode], C.c_void_p(stream or 0), _b(major), _b(ta), _b(tb), M, N, K, alpha,
                              _ptr(dA), lda, _ptr(dB), ldb, beta, _ptr(dC), ldc)
    if rc:
        check()
        raise UgemmCudaError("sgemm_cuda_dev failed")


def sgemm_cuda_batched(major, ta, tb, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, Cm, ldc, strideC, batch):
    """`batch` stacked instances in one launch (host buffers, blocking)."""
    lib().sgemm_cuda_batched(_b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(A), lda, strideA, _ptr(B), ldb, strideB, beta,
                             _ptr(Cm), ldc, strideC, batch)
    check()


def sgemm_cuda_batched_dev(mode, stream, major, ta, tb, M, N, K, alpha, dA, lda, strideA, dB, ldb, strideB, beta, dC, ldc, strideC, batch):
    rc = lib().sgemm_cuda_batched_dev(_MODES[mode], C.c_void_p(stream or 0), _b(major), _b(ta), _b(tb), M, N, K, alpha,
                                      _ptr(dA), lda, strideA, _ptr(dB), ldb, strideB, beta, _ptr(dC), ldc, strideC, batch)
    if rc:
        check()
        raise UgemmCudaError("sgemm_cuda_batched_dev failed")


def sgemm_cuda_time_dev(mode, iters, warmup, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, total=False):
    """(avg_ms, min_ms[, total_ms]) of `iters` back-to-back launches timed with CUDA events on the backend stream."""
    avg, best, span = C.c_float(0), C.c_float(0), C.c_float(0)
    rc = lib().sgemm_cuda_time_dev(_MODES[mode], iters, warmup, _b(major), _b(ta), _b(tb), M, N, K, alpha,
                                   _ptr(dA), lda, _ptr(dB), ldb, beta, _ptr(dC), ldc, C.byref(avg), C.byref(best), C.byref(span))
    if rc:
        check()
        raise UgemmCudaError("sgemm_cuda_time_dev failed")
    return (avg.value, best.value, span.value) if total else (avg.value, best.value)


def k1_eligible(major, ta, tb, M, N, K, dA, lda, dB, ldb, dC, ldc):
    return bool(lib().sgemm_cuda_k1_eligible(_b(major), _b(ta), _b(tb), M, N, K, _ptr(dA), lda, _ptr(dB), ldb, _ptr(dC), ldc))


def set_k1_tuning(kc_blocks=-1, split=-1, cta_group=-1):
    lib().sgemm_cuda_set_k1_tuning(kc_blocks, split, cta_group)


K1_PLAN_KEYS = ("cta_group", "tile_m", "tile_n", "tiles_m", "tiles_n", "k_blocks", "kc", "whole_tiles", "tail_tiles", "chunks_per_tile", "chunks_per_range", "items")


def k1_plan(M, N, K, batch=1, sm_count=0):
    """The schedule K1 would use (pure host arithmetic, no GPU): dict of K1_PLAN_KEYS plus the raw int array under "_raw"."""
    raw = (C.c_int * 12)()
    if lib().sgemm_cuda_k1_plan(M, N, K, batch, sm_count, raw):
        check()
        raise UgemmCudaError("sgemm_cuda_k1_plan failed")
    d = dict(zip(K1_PLAN_KEYS, list(raw)))
    d["_raw"] = raw
    return d


def k1_plan_item(plan, item, h):
    """(tile, kb0, kb1, slot) of segment h of work item `item` of a k1_plan()"""
    out = (C.c_int * 4)()
    if lib().sgemm_cuda_k1_plan_item(plan["_raw"], item, h, out):
        check()
        raise UgemmCudaError("sgemm_cuda_k1_plan_item failed")
    return tuple(out)


def set_k1_variant(variant=0):
    """0 = TS kernel (A in tensor memory, default), 1 = round-1 SS kernel."""
    lib().sgemm_cuda_set_k1_variant(int(variant))


def set_sm_limit(sms=0):
    lib().sgemm_cuda_set_sm_limit(int(sms))


def last_kernel():
    return {0: None, 1: "3xtf32", 2: "simt"}[lib().sgemm_cuda_last_kernel()]


def last_repacked():
    return bool(lib().sgemm_cuda_last_repacked())


def launch_count():
    return int(lib().sgemm_cuda_launch_count())


def device_info():
    sm, khz, mem = C.c_int(0), C.c_int(0), C.c_size_t(0)
    name = C.create_string_buffer(128)
    if lib().ugemm_cuda_device_info(C.byref(sm), C.byref(khz), C.byref(mem), name, 128):
        check()
    return {"sm_count": sm.value, "sm_clock_khz": khz.value, "hbm_bytes": mem.value, "name": name.value.decode()}


class DeviceBuffer:
    """A float32 device allocation owned through the C ABI (ugemm_cuda_malloc / ugemm_cuda_free)."""

    def __init__(self, n_floats):
        self.n = int(n_floats)
        self.ptr = lib().ugemm_cuda_malloc(max(self.n, 1) * 4)
        if not self.ptr:
            check()
            raise UgemmCudaError("ugemm_cuda_malloc failed")

    def data_ptr(self):
        return self.ptr

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float32)
        assert host.size <= self.n
        if lib().ugemm_cuda_memcpy_h2d(C.c_void_p(self.ptr), _ptr(host), host.size * 4):
            check()
        return self

    def download(self, n=None, offset=0):
        n = self.n - offset if n is None else n
        out = np.empty(n, dtype=np.float32)
        if lib().ugemm_cuda_memcpy_d2h(_ptr(out), C.c_void_p(self.ptr + offset * 4), n * 4):
            check()
        return out

    def fill_uniform(self, seed, lo=0.0, hi=1.0, n=None, offset=0):
        n = self.n - offset if n is None else n
        if lib().ugemm_fill_uniform_dev(C.c_void_p(self.ptr + offset * 4), n, seed, lo, hi, None):
            check()
        return self

    def free(self):
        if self.ptr:
            lib().ugemm_cuda_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fill_uniform_host(n, seed, lo=0.0, hi=1.0):
    x = np.empty(int(n), dtype=np.float32)
    lib().ugemm_fill_uniform_host(_ptr(x), x.size, seed, lo, hi)
    return x


def fill_uniform_host_2d(rows, cols, seed, offset, gld, lo=0.0, hi=1.0, ld=None, out=None):
    """rows x cols window (pitch ld) of stream `seed` starting at flat index `offset` of a matrix with pitch gld."""
    ld = cols if ld is None else ld
    x = np.empty(int(rows) * int(ld), dtype=np.float32) if out is None else out
    lib().ugemm_fill_uniform_host_2d(_ptr(x), rows, cols, ld, seed, offset, gld, lo, hi)
    return x


def fill_uniform_dev_2d(dptr, rows, cols, ld, seed, offset, gld, lo=0.0, hi=1.0, stream=None):
    if lib().ugemm_fill_uniform_dev_2d(_ptr(dptr), rows, cols, ld, seed, offset, gld, lo, hi, C.c_void_p(stream or 0)):
        check()


def memcpy_async(dst, src, nbytes, stream=None):
    if lib().ugemm_cuda_memcpy_async(_ptr(dst), _ptr(src), nbytes, C.c_void_p(stream or 0)):
        check()


def ipc_export(dptr):
    """64-byte CUDA IPC handle (bytes) of a buffer allocated with ugemm_cuda_malloc / DeviceBuffer."""
    h = C.create_string_buffer(64)
    if lib().ugemm_cuda_ipc_export(_ptr(dptr), h):
        check()
    return h.raw


def ipc_import(handle):
    p = lib().ugemm_cuda_ipc_import(C.c_char_p(handle))
    if not p:
        check()
        raise UgemmCudaError("ugemm_cuda_ipc_import failed")
    return p


def ipc_close(ptr):
    lib().ugemm_cuda_ipc_close(C.c_void_p(ptr))


class Shard:
    """ctypes mirror of the one-process-per-GPU sharded SGEMM of the C ABI (csrc/shard.cu, sgemm_cuda_shard_*).  The caller brings
    rendezvous only: rank 0 calls Shard.unique_id() and hands the 128 bytes to every rank."""
    NCCL, P2P = 0, 1

    @staticmethod
    def plan(world, rank, M, N, K):
        v = [C.c_int(0) for _ in range(6)]
        if lib().sgemm_cuda_shard_plan(world, rank, M, N, K, *[C.byref(x) for x in v]):
            check()
            raise UgemmCudaError("sgemm_cuda_shard_plan failed")
        return dict(zip(("pr", "pc", "L", "kw", "mloc", "nloc"), (x.value for x in v)))

    @staticmethod
    def owners(world, rank, M, N, K, t):
        ao, bo, aoff, boff = C.c_int(0), C.c_int(0), C.c_longlong(0), C.c_longlong(0)
        if lib().sgemm_cuda_shard_owners(world, rank, M, N, K, t, C.byref(ao), C.byref(bo), C.byref(aoff), C.byref(boff)):
            check()
            raise UgemmCudaError("sgemm_cuda_shard_owners failed")
        return ao.value, bo.value, aoff.value, boff.value

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        if lib().sgemm_cuda_shard_unique_id(buf):
            check()
            raise UgemmCudaError("sgemm_cuda_shard_unique_id failed")
        return buf.raw

    def __init__(self, rank, world, uid, M, N, K, transport=1):
        self.rank, self.world, self.M, self.N, self.K = rank, world, M, N, K
        if lib().sgemm_cuda_shard_init(rank, world, uid, M, N, K, transport):
            check()
            raise UgemmCudaError("sgemm_cuda_shard_init failed")
        self.transport = lib().sgemm_cuda_shard_transport()
        self.p = Shard.plan(world, rank, M, N, K)

    def _call(self, name, *args):
        if getattr(lib(), "sgemm_cuda_shard_" + name)(*args):
            check()
            raise UgemmCudaError(f"sgemm_cuda_shard_{name} failed")

    def generate(self, seed_a=1, seed_b=2, lo=0.0, hi=1.0):
        self._call("generate", seed_a, seed_b, lo, hi)

    def run(self, distribute=True, steps=1, warmup=0):
        """this rank's CUDA-event milliseconds for `steps` steps (take the max over ranks with allreduce_max)"""
        ms = C.c_float(0)
        self._call("run", 1 if distribute else 0, steps, warmup, C.byref(ms))
        return ms.value

    def run_host(self, steps=1, warmup=0):
        ms, up, down = C.c_float(0), C.c_longlong(0), C.c_longlong(0)
        self._call("run_host", steps, warmup, C.byref(ms), C.byref(up), C.byref(down))
        return ms.value, up.value, down.value

    def download_owned(self):
        self._call("download_owned")

    def copy_floor(self, steps=2):
        """host wall-clock ms of `steps` x (owned slabs up + C block down at once, nothing else): the host-link floor of run_host"""
        ms = C.c_float(0)
        self._call("copy_floor", steps, C.byref(ms))
        return ms.value

    def allreduce(self, value, op="max"):
        v = C.c_float(value)
        self._call("allreduce", C.byref(v), 0 if op == "max" else 1)
        return v.value

    def block(self):
        """(device pointer, rows, cols, row0, col0) of this rank's C block"""
        ptr, r, c, r0, c0 = C.c_void_p(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._call("block", C.byref(ptr), C.byref(r), C.byref(c), C.byref(r0), C.byref(c0))
        return ptr.value, r.value, c.value, r0.value, c0.value

    def host_c(self):
        import numpy as np
        hown, n1, hc, n2 = C.c_void_p(0), C.c_longlong(0), C.c_void_p(0), C.c_longlong(0)
        self._call("host_buffers", C.byref(hown), C.byref(n1), C.byref(hc), C.byref(n2))
        return np.ctypeslib.as_array(C.cast(hc, C.POINTER(C.c_float)), shape=(n2.value,))

    def finish(self):
        lib().sgemm_cuda_shard_finish()


def sgemm_cuda_mgpu_init(ngpus):
    """GPUs 0..ngpus-1 driven by this one process (peer access, per-GPU streams and arenas)."""
    if lib().sgemm_cuda_mgpu_init(int(ngpus)):
        check()
        raise UgemmCudaError("sgemm_cuda_mgpu_init failed")


def sgemm_cuda_mgpu_finish():
    lib().sgemm_cuda_mgpu_finish()


def sgemm_cuda_mgpu_count():
    return int(lib().sgemm_cuda_mgpu_count())


def sgemm_cuda_mgpu_plan(M, N, K, pr, pc, overlap=1):
    """(block_rows, block_cols, k_slabs, slab_width) of sgemm_cuda_mgpu's partition; host arithmetic only."""
    v = [C.c_int(0) for _ in range(4)]
    if lib().sgemm_cuda_mgpu_plan(M, N, K, pr, pc, overlap, *[C.byref(x) for x in v]):
        raise UgemmCudaError("sgemm_cuda_mgpu_plan: bad arguments")
    return tuple(x.value for x in v)


def visible_gpus():
    """Number of CUDA devices this process can see (0 without a driver)."""
    return int(lib().ugemm_cuda_device_count())


def sgemm_cuda_mgpu(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, pr, pc, overlap=1):
    """Sharded SGEMM over a pr x pc grid of GPUs from ONE process; A/B/C host arrays or device pointers of any GPU.
    Returns the 5 timings (ms): host wall, device span, distribution span, product span, span without the C write-back."""
    t = (C.c_float * 5)()
    rc = lib().sgemm_cuda_mgpu(_b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(A), lda, _ptr(B), ldb, beta, _ptr(Cm), ldc,
                               int(pr), int(pc), int(overlap), C.cast(t, C.c_void_p))
    if rc:
        check()
        raise UgemmCudaError("sgemm_cuda_mgpu failed")
    return tuple(t)


def sync():
    if lib().ugemm_cuda_sync():
        check()


def probe_tf32(A, B, ksteps):
    """A: (128, 8*ksteps), B: (16, 8*ksteps) float32 -> D (128, 16): chained TF32 tcgen05 MMAs."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    assert A.shape == (128, 8 * ksteps) and B.shape == (16, 8 * ksteps)
    D = np.empty((128, 16), dtype=np.float32)
    if lib().ugemm_cuda_probe_tf32(_ptr(A), _ptr(B), _ptr(D), ksteps):
        check()
        raise UgemmCudaError("probe failed")
    return D


# ---- convolution callers (argument order of ocl_convolution, sgemm_ocl1.h:271, and gl_convolution_LReLU, sgemm_gl1.h:192)
def im2col_cuda(im, channels, height, width, k, pad, stride, col):
    lib().im2col_cuda(_ptr(im), channels, height, width, k, pad, stride, _ptr(col))
    check()


def convolution_cuda(inputs, ich, w, h, weights, k, pad, stride, outputs, ch):
    lib().convolution_cuda(_ptr(inputs), ich, w, h, _ptr(weights), k, pad, stride, _ptr(outputs), ch)
    check()


def convolution_cuda_LReLU(inputs, ich, w, h, weights, k, pad, stride, outputs, ch, bias):
    lib().convolution_cuda_LReLU(_ptr(inputs), ich, w, h, _ptr(weights), k, pad, stride, _ptr(outputs), ch, _ptr(bias))
    check()


def convolution_cuda_dev(mode, stream, d_inputs, ich, w, h, d_weights, k, pad, stride, d_outputs, ch, d_bias, slope, d_workspace):
    rc = lib().convolution_cuda_dev(_MODES[mode], C.c_void_p(stream or 0), _ptr(d_inputs), ich, w, h, _ptr(d_weights), k, pad, stride,
                                    _ptr(d_outputs), ch, _ptr(d_bias), slope, _ptr(d_workspace))
    if rc:
        check()
        raise UgemmCudaError("convolution_cuda_dev failed")


def convolution_cuda_batched_dev(mode, stream, d_inputs, nimg, ich, w, h, d_weights, k, pad, stride, d_outputs, ch, d_bias, slope, d_workspace):
    """nimg images [nimg][ich][h][w] -> [nimg][ch][Ho*Wo]; one fused implicit-GEMM launch when the layout allows."""
    rc = lib().convolution_cuda_batched_dev(_MODES[mode], C.c_void_p(stream or 0), _ptr(d_inputs), nimg, ich, w, h, _ptr(d_weights), k, pad, stride,
                                            _ptr(d_outputs), ch, _ptr(d_bias), slope, _ptr(d_workspace))
    if rc:
        check()
        raise UgemmCudaError("convolution_cuda_batched_dev failed")


def set_conv_fusion(mode=-1):
    """-1 by rule, 0 never (im2col + GEMM), 1 implicit GEMM whenever the hard constraints allow."""
    lib().sgemm_cuda_set_conv_fusion(int(mode))


def last_conv_fused():
    return bool(lib().sgemm_cuda_last_conv_fused())


# ---- level 1 / level 2 (argument order of saxpy_cpu ugemm.h:75 and sgemv_cpu ugemm.h:124) ----------------
def saxpy_cuda(N, alpha, x, incx, y, incy):
    lib().saxpy_cuda(N, alpha, _ptr(x), incx, _ptr(y), incy)
    check()


def saxpy_cuda_dev(stream, N, alpha, dx, incx, dy, incy):
    if lib().saxpy_cuda_dev(C.c_void_p(stream or 0), N, alpha, _ptr(dx), incx, _ptr(dy), incy):
        check()
        raise UgemmCudaError("saxpy_cuda_dev failed")


def sgemv_cuda(trans, M, N, alpha, A, lda, x, incx, beta, y, incy):
    lib().sgemv_cuda(_b(trans), M, N, alpha, _ptr(A), lda, _ptr(x), incx, beta, _ptr(y), incy)
    check()


def sgemv_cuda_dev(stream, trans, M, N, alpha, dA, lda, dx, incx, beta, dy, incy):
    if lib().sgemv_cuda_dev(C.c_void_p(stream or 0), _b(trans), M, N, alpha, _ptr(dA), lda, _ptr(dx), incx, beta, _ptr(dy), incy):
        check()
        raise UgemmCudaError("sgemv_cuda_dev failed")


# ---- DGEMM (the uut signature of check_dgemm.c:86-97) ------------------------------------------------------
def dgemm_cuda(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    """Host float64 buffers, blocking, C updated in place."""
    f8 = np.float64
    lib().dgemm_cuda(_b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(A, f8), lda, _ptr(B, f8), ldb, beta, _ptr(Cm, f8), ldc)
    check()


def dgemm_cuda_dev(stream, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc):
    if lib().dgemm_cuda_dev(C.c_void_p(stream or 0), _b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(dA), lda, _ptr(dB), ldb, beta, _ptr(dC), ldc):
        check()
        raise UgemmCudaError("dgemm_cuda_dev failed")


def dgemm_cuda_time_dev(iters, warmup, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc):
    avg, best = C.c_float(0), C.c_float(0)
    if lib().dgemm_cuda_time_dev(iters, warmup, _b(major), _b(ta), _b(tb), M, N, K, alpha, _ptr(dA), lda, _ptr(dB), ldb, beta, _ptr(dC), ldc,
                                 C.byref(avg), C.byref(best)):
        check()
        raise UgemmCudaError("dgemm_cuda_time_dev failed")
    return avg.value, best.value


# ---- the macro API of the reference's GPU harness (sgemm_test.c:19-33): tight row-major, no ld ----------
def sgemm_init(s1=0, s2=0, s3=0):
    """sgemm_init(M*K, K*N, M*N) -> sgemm_ocl_init(0, 0, (s1+s2+s3)*10*sizeof(float)) in the reference."""
    sgemm_cuda_init(-1, (s1 + s2 + s3) * 4)


def sgemm_finish():
    sgemm_cuda_finish()


def sgemm_rnn(M, N, K, alpha, a, b, beta, c):
    sgemm_cuda("R", "N", "N", M, N, K, alpha, a, K, b, N, beta, c, N)


def sgemm_rnt(M, N, K, alpha, a, b, beta, c):
    sgemm_cuda("R", "N", "T", M, N, K, alpha, a, K, b, K, beta, c, N)


def sgemm_rtn(M, N, K, alpha, a, b, beta, c):
    sgemm_cuda("R", "T", "N", M, N, K, alpha, a, M, b, N, beta, c, N)
