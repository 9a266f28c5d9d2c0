/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin export shim around the UNMODIFIED reference headers.  The reference
 * sources are NOT copied into this repository: this file is compiled with
 * -I$UGEMM_REF (default /root/reference) by oracle/Makefile and only the
 * resulting shared object lands in oracle/_ref/ (git-ignored, travels to the
 * GPU box with the gpurun snapshot).
 *
 * Exported entry points (all have the reference's 14-argument BLAS signature,
 * check_sgemm.c:96-103):
 *   ref_sgemm_cpu  -> sgemm_cpu   ugemm.h:287       naive ground truth
 *   ref_sgemm_c    -> sgemm_c     gemm_cpu.h:284    Goto-blocked scalar 4x4
 *   ref_sgemm_avx  -> sgemm_avx   sgemm_avx256.h:392  AVX "noncblas", row-major NN only
 *   ref_sgemm_sse  -> sgemm_sse   sgemm_sse.h:365   AVX 8x8 Goto (static in the header)
 *   ref_saxpy_cpu  -> saxpy_cpu   ugemm.h:75        y += alpha*x
 *   ref_sgemv_cpu  -> sgemv_cpu   ugemm.h:124       naive gemv
 *   ref_dgemm_cpu  -> dgemm_cpu   ugemm.h:162       naive double-precision ground truth
 *   ref_dgemm_c    -> _dgemm_c    gemm_cpu.h:284 with real = double (ugemm.h:29-33)
 *   ref_dgemm_avx  -> dgemm_avx   dgemm_avx.h:844   AVX Goto DGEMM, all majors / transposes
 *   ref_sgemm_avx_mt: harness-level wrapper, sgemm_avx on disjoint row slabs of
 *                  A/C from `threads` OpenMP threads.  Legal because all state of
 *                  avx256_noncblas_sgemm lives in the stack-allocated
 *                  noncblas_sgemm_prm_t (sgemm_avx256.h:334).  NOT reference code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "ugemm.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define SIG char major, char ta, char tb, int M, int N, int K, float alpha, \
            const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc
#define ARGS major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc

void ref_sgemm_cpu(SIG) { sgemm_cpu(ARGS); }
void ref_sgemm_c  (SIG) { sgemm_c(ARGS); }
void ref_sgemm_avx(SIG) { sgemm_avx(ARGS); }
void ref_sgemm_sse(SIG) { sgemm_sse(ARGS); }

void ref_saxpy_cpu(int N, float alpha, const float *x, int incx, float *y, int incy) { saxpy_cpu(N, alpha, x, incx, y, incy); }
void ref_sgemv_cpu(char trans, int M, int N, float alpha, const float *A, int lda, const float *x, int incx,
                   float beta, float *y, int incy)
{ sgemv_cpu(trans, M, N, alpha, A, lda, x, incx, beta, y, incy); }

#define DSIG char major, char ta, char tb, int M, int N, int K, double alpha, \
             const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc
void ref_dgemm_cpu(DSIG) { dgemm_cpu(ARGS); }
void ref_dgemm_c  (DSIG) { _dgemm_c(ARGS); }
void ref_dgemm_avx(DSIG) { dgemm_avx(ARGS); }

int ref_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* row-major NN only (that is all sgemm_avx implements). Slabs are multiples of
 * 2 rows because the AVX core walks M in steps of 2 (sgemm_avx256.h:28). */
void ref_sgemm_avx_mt(int threads, SIG)
{
	if (threads <= 1 || M < 4 * threads) { sgemm_avx(ARGS); return; }
	int slab = ((M + threads - 1) / threads + 1) & ~1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static, 1)
#endif
	for (int t = 0; t < threads; t++) {
		int m0 = t * slab;
		int mm = M - m0 < slab ? M - m0 : slab;
		if (mm > 0)
			sgemm_avx(major, ta, tb, mm, N, K, alpha,
			          A + (size_t)m0 * lda, lda, B, ldb, beta,
			          C + (size_t)m0 * ldc, ldc);
	}
}
