// common.cuh -- shared declarations of the CUDA backend (internal; the public ABI is include/ugemm_cuda.h)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace ugemm {

// A GEMM already normalised to row-major (column-major callers are mapped by operand swap, see
// backend.cu: normalise()).  a_kmajor: op(A)'s k index is the contiguous one in memory (transA=='N');
// b_kmajor: op(B)'s k index is contiguous (transB=='T').
struct Problem {
	int M, N, K;
	float alpha, beta;
	const float *A; long long lda; bool a_kmajor;
	const float *B; long long ldb; bool b_kmajor;
	float *C; long long ldc;
	// optional fused epilogue of the convolution callers (sgemm_gl1.h:210-217): out = act(alpha*acc + beta*C + bias[m]),
	// act(x) = x > 0 ? x : slope*x.  bias == nullptr and slope == 1 (the default) mean a plain GEMM.
	const float *bias = nullptr;
	float slope = 1.f;
	// strided batch (the stacked-instance layout of test_sgemm, check_sgemm.c:116-120): instance b uses
	// A + b*strideA, B + b*strideB, C + b*strideC (elements); batch == 1 is a plain GEMM.
	int batch = 1;
	long long strideA = 0, strideB = 0, strideC = 0;
};

// DGEMM (K4) problem, normalised to row-major like Problem
struct DProblem {
	int M, N, K;
	double alpha, beta;
	const double *A; long long lda; bool a_kmajor;
	const double *B; long long ldb; bool b_kmajor;
	double *C; long long ldc;
};

// Implicit-GEMM convolution on K1: out[img][co][io*wo + jo] = sum_{c,ki,kj} w[co][c][ki][kj] * in[img][c][io*stride+ki-pad][jo*stride+kj-pad]
// (+ bias[co], LeakyReLU).  The column matrix of the reference's im2col (sgemm_ocl1.h:81-119) is never built: the B operand
// tiles are gathered by 4-D TMA boxes from in_hwc, a channels-last copy [img][y][x][cs] of the image (one image-sized pass,
// launch_chw_to_hwc; cs = ich rounded up to 4).  wgt_kkc = weights repacked to [co][ki*k+kj][ichp], ichp = ich rounded up to 32.
struct ConvProblem {
	const float *in_hwc; int cs; int nimg, ich, h, w;
	const float *wgt_kkc; int ichp;
	int k, pad, stride, ho, wo, ch;
	float *out; const float *bias; float slope;
};

// flags: bit0 = share one shared-memory read of A_big between big*small and big*big (A collector);
// bits 1-4 are ABLATION switches for bottleneck analysis only (results are wrong with them):
// 2 transform skips its stores, 4 transform skips loads and stores, 8 only big*big is issued, 16 epilogue skips stores,
// 32 per-role cycle counters to stderr, 64 MMA free-run (no TMA / transform / stage barriers: raw tensor-pipe ceiling).
// bits 7-10: L2 eviction hints of the TMA loads (experiments); bit 11 (2048): stream-K tail off (A/B runs, SM-limited launches).
// bit 13 (8192): TMA-store epilogue off (row-strided 128-bit global stores instead; A/B runs).
// bit 15 (32768): round-1 SS kernel (A and B from shared memory) instead of the TS kernel (A in tensor memory); A/B runs.
// bit 19 (524288): serpentine K off (every tile walks its k-blocks upwards; A/B runs of the DRAM traffic).
// bit 18 (262144): programmatic dependent launch off (A/B runs).
// bit 17 (131072): stream-K tail whenever the model predicts any saving (A/B runs of the rule).
// bit 16 (65536): ABLATION, stream-K fix-up pass skipped (results wrong).
// bit 20 (1048576): register path of the beta != 0 preload with 128-bit loads instead of 256-bit ones (A/B runs).
// bit 22 (4194304): beta != 0 takes the old C into the registers with global loads when a group is re-armed (the path of a C the TMA
// unit cannot address) instead of fetching it box by box through the TMA unit and adding it between promotions (A/B runs).
// bit 14 (16384): round-1 barrier arrives (.release.cluster = MEMBAR.ALL.GPU in front of every transform / epilogue arrive; A/B runs).
struct K1Tuning { int kc_blocks; int split; int cta_group; int flags; };

// K2: register-blocked FFMA kernel (k2_simt.cu)
cudaError_t launch_k2_simt(const Problem &p, cudaStream_t stream, int sm_count);
// K1: 3xTF32 tcgen05 kernel (k1_tcgen05.cu).  *why (optional) receives a static string on ineligibility.
bool        k1_eligible(const Problem &p, const char **why);
cudaError_t launch_k1_3xtf32(const Problem &p, const K1Tuning &t, cudaStream_t stream, int sm_count);
// the schedule of a dense K1 launch as host arithmetic (k1_tcgen05.cu; exported as sgemm_cuda_k1_plan / _plan_item)
void k1_plan(int M, int N, int K, int batch, const K1Tuning &t, int sm_count, int *out12);
void k1_plan_item(int item, int h, int sk_full, int sk_rem, int sk_nch, int sk_q, int kc, int nkb, int *out4);
// implicit-GEMM convolution (k1_tcgen05.cu): weight repack [co][c][ki][kj] -> [co][ki*k+kj][ichp] (zero padded), and the launch
cudaError_t launch_conv_weight_repack(const float *w, int ch, int ich, int k, int ichp, float *dst, cudaStream_t stream);
cudaError_t launch_chw_to_hwc(const float *in, int nimg, int ich, int h, int w, int cs, float *out, cudaStream_t stream);
cudaError_t launch_k1_conv(const ConvProblem &c, const K1Tuning &t, cudaStream_t stream, int sm_count);
// C <- beta*C over the M x N region (alpha==0 or K==0 path)
cudaError_t launch_scale_c(const Problem &p, cudaStream_t stream);
// im2col of a planar C x H x W image into the (C*k*k) x (Ho*Wo) column matrix (k2_simt.cu)
cudaError_t launch_im2col(const float *im, int channels, int height, int width, int k, int pad, int stride, float *col, cudaStream_t stream);
// level 1 / level 2 companions (k3_level12.cu): y += alpha*x;  y = alpha*op(A)*x + beta*y
cudaError_t launch_saxpy(long long n, float alpha, const float *x, long long incx, float *y, long long incy, cudaStream_t stream, int sm_count);
cudaError_t launch_sgemv(bool rows_contiguous, int M, int N, float alpha, const float *A, long long lda, const float *x, long long incx,
                         float beta, float *y, long long incy, cudaStream_t stream, int sm_count);
// K4: register-blocked DFMA DGEMM (k4_dgemm.cu) and its C <- beta*C companion
cudaError_t launch_k4_dgemm(const DProblem &p, cudaStream_t stream);
cudaError_t launch_scale_c_f64(const DProblem &p, cudaStream_t stream);
// probe (k1_tcgen05.cu)
cudaError_t launch_probe_tf32(const float *dA, const float *dB, float *dD, int ksteps, cudaStream_t stream);
// last diagnostic record written by a K1 watchdog (host-mapped memory), 0 if none
const unsigned *k1_diag_host();

// counter-based uniform stream shared by host and device (see ugemm_cuda.h)
__host__ __device__ inline unsigned long long mix64(unsigned long long z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__host__ __device__ inline float uniform_at(unsigned long long base, unsigned long long i, float lo, float span)
{
	float u = (float)(mix64(base + i) >> 40) * 5.9604644775390625e-08f; // 2^-24, exact
	return fmaf(span, u, lo);
}

} // namespace ugemm
