"""ugemm_b200 -- Blackwell (sm_100a) SGEMM backend for ugemm: Python host mirror of the C ABI.

Only what the hot path needs: `csrc/` (hand-written CUDA kernels + the extern "C" boundary declared in
include/ugemm_cuda.h), `build.py` (nvcc recipe) and `backend.py` (ctypes binding that mirrors the
reference's `sgemm_*` entry points), plus `dist.py` (2-D C-tile sharding across GPUs).
"""
from . import backend  # noqa: F401
from .backend import (  # noqa: F401
    MODE_3XTF32, MODE_AUTO, MODE_SIMT, convolution_cuda, convolution_cuda_dev, convolution_cuda_batched_dev, set_conv_fusion, last_conv_fused, convolution_cuda_LReLU, im2col_cuda, DeviceBuffer, UgemmCudaError, check, device_info, fill_uniform_host,
    fill_uniform_dev_2d, fill_uniform_host_2d,
    k1_eligible, k1_plan, k1_plan_item, last_error, last_kernel, last_repacked, launch_count, lib, probe_tf32, set_k1_tuning, set_k1_variant, set_sm_limit, sgemm_cuda,
    sgemm_cuda_3xtf32, sgemm_cuda_batched, sgemm_cuda_batched_dev, sgemm_cuda_dev, sgemm_cuda_finish, sgemm_cuda_init, sgemm_cuda_simt, sgemm_cuda_time_dev,
    sgemm_finish, sgemm_init, sgemm_rnn, sgemm_rnt, sgemm_rtn, sync,
    Shard, sgemm_cuda_mgpu, sgemm_cuda_mgpu_count, sgemm_cuda_mgpu_finish, sgemm_cuda_mgpu_init, sgemm_cuda_mgpu_plan, visible_gpus,
    saxpy_cuda, saxpy_cuda_dev, sgemv_cuda, sgemv_cuda_dev, dgemm_cuda, dgemm_cuda_dev, dgemm_cuda_time_dev,
)
