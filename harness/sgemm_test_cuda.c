/* sgemm_test_cuda.c -- the reference's GPU-backend harness (sgemm_test.c) with a CUDA branch.
 *
 * sgemm_test.c selects a backend at compile time and talks to it through five macros (sgemm_test.c:9-34):
 *     sgemm_init(s1,s2,s3)  sgemm_finish()  sgemm_rnn(...)  sgemm_rnt(...)  sgemm_rtn(...)
 * This file is that harness with the third `#elif CATS_CUDA` branch SURVEY.md §3.5 describes, kept otherwise in the
 * reference's shape: 1023 x 1000 x 1023 (not tile multiples, sgemm_test.c:168-172), ramp inputs A[i]=B[i]=i+1
 * (sgemm_test.c:219-220; entries reach 1e15..1e18 so the comparison is purely relative), 20 repetitions timed with
 * the PCIe copies inside (sgemm_test.c:229-232), CPU loops gemm_rnn / gemm_rnt / gemm_rtn as the reference
 * (sgemm_test.c:40-104, restated below with 64-bit indexing), the cmp_results line (sgemm_test.c:106-135) and the
 * ">>> Done" line (sgemm_test.c:164) with the integer-division GFLOP bug fixed.  Also runs the 3x2 . 2x3 known-answer
 * case from the comment at sgemm_test.c:186-200.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CATS_CUDA
#ifdef CATS_CUDA
#include "ugemm_cuda.h"
#define sgemm_init(s1, s2, s3)  sgemm_cuda_init(-1, (size_t)((s1) + (s2) + (s3)) * sizeof(float))
#define sgemm_finish()          sgemm_cuda_finish()
#define sgemm_rnn(m, n, k, alpha, a, b, beta, c) sgemm_cuda('R', 'N', 'N', m, n, k, alpha, a, k, b, n, beta, c, n)
#define sgemm_rnt(m, n, k, alpha, a, b, beta, c) sgemm_cuda('R', 'N', 'T', m, n, k, alpha, a, k, b, k, beta, c, n)
#define sgemm_rtn(m, n, k, alpha, a, b, beta, c) sgemm_cuda('R', 'T', 'N', m, n, k, alpha, a, m, b, n, beta, c, n)
#endif

#define MSIZE 1023
#define NSIZE 1000
#define KSIZE 1023
#define REPS 20

/* CPU references: plain loops in the reference's index conventions (sgemm_test.c:40-104) */
static void gemm_rnn(int M, int N, int K, float alpha, const float *A, const float *B, float beta, float *C)
{
	for (int m = 0; m < M; m++)
		for (int n = 0; n < N; n++) {
			float acc = 0;
			for (int k = 0; k < K; k++) acc += A[(size_t)m * K + k] * B[(size_t)k * N + n];
			C[(size_t)m * N + n] = alpha * acc + beta * C[(size_t)m * N + n];
		}
}
static void gemm_rnt(int M, int N, int K, float alpha, const float *A, const float *B, float beta, float *C)
{
	for (int m = 0; m < M; m++)
		for (int n = 0; n < N; n++) {
			float acc = 0;
			for (int k = 0; k < K; k++) acc += A[(size_t)m * K + k] * B[(size_t)n * K + k];   /* B is N x K, sgemm_test.c:76 */
			C[(size_t)m * N + n] = alpha * acc + beta * C[(size_t)m * N + n];
		}
}
static void gemm_rtn(int M, int N, int K, float alpha, const float *A, const float *B, float beta, float *C)
{
	for (int m = 0; m < M; m++)
		for (int n = 0; n < N; n++) {
			float acc = 0;
			for (int k = 0; k < K; k++) acc += A[(size_t)k * M + m] * B[(size_t)k * N + n];   /* A is K x M, sgemm_test.c:98 */
			C[(size_t)m * N + n] = alpha * acc + beta * C[(size_t)m * N + n];
		}
}

static double cmp_results(int M, int N, const float *ref, const float *res, int ld)
{
	double maxErr = 0, s2Err = 0, s1Ref = 0, s2Ref = 0;
	size_t maxI = 0;
	for (int m = 0; m < M; ++m)
		for (int n = 0; n < N; ++n) {
			double refV = ref[(size_t)m * ld + n], resV = res[(size_t)m * ld + n], err = resV - refV;
			if (maxErr < fabs(err)) { maxErr = fabs(err); maxI = (size_t)m * ld + n; }
			s2Err += err * err; s1Ref += refV; s2Ref += refV * refV;
		}
	double mn = (double)M * N;
	double stdErr = sqrt(s2Err / mn), stdRef = sqrt(s2Ref * mn - s1Ref * s1Ref) / mn;
	printf("%.3e/%.3e=%.3e. %.3e at [%3zu,%3zu] %18.10e vs %18.10e %s\n", stdErr, stdRef, stdErr / stdRef, maxErr, maxI / ld,
	       maxI % ld, (double)ref[maxI], (double)res[maxI],
	       maxErr > stdRef * 1e-5 ? "FAIL !!!" : (maxErr > stdRef * 3e-5 || stdErr > stdRef * 1e-6 ? "Sucks !" : ""));
	return s2Ref > 0 ? sqrt(s2Err / s2Ref) : 0;
}

static double t_start;
static void start(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); t_start = ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void end(int reps)
{
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	double runtime = (ts.tv_sec + 1e-9 * ts.tv_nsec - t_start) / reps;
	double gflop = 2.0 * MSIZE * NSIZE * KSIZE / 1e9;   /* the reference divides integers here and always gets 2 */
	printf(">>> Done: took %.3lf seconds per run, %.1lf GFLOPS\n", runtime, gflop / runtime);
}

static float A[MSIZE * KSIZE], B[KSIZE * NSIZE], C[MSIZE * NSIZE], Z[MSIZE * NSIZE];

static int run_case(const char *name, int which)
{
	for (int i = 0; i < MSIZE * NSIZE; i++) C[i] = Z[i] = 1;
	start();
	for (int r = 0; r < REPS; r++) {
		if (which == 0) sgemm_rnn(MSIZE, NSIZE, KSIZE, 1, A, B, 0, C);
		if (which == 1) sgemm_rnt(MSIZE, NSIZE, KSIZE, 1, A, B, 0, C);
		if (which == 2) sgemm_rtn(MSIZE, NSIZE, KSIZE, 1, A, B, 0, C);
	}
	end(REPS);
	if (sgemm_cuda_last_error()) { printf("%s: %s\n", name, sgemm_cuda_last_error()); return 1; }
	if (which == 0) gemm_rnn(MSIZE, NSIZE, KSIZE, 1, A, B, 0, Z);
	if (which == 1) gemm_rnt(MSIZE, NSIZE, KSIZE, 1, A, B, 0, Z);
	if (which == 2) gemm_rtn(MSIZE, NSIZE, KSIZE, 1, A, B, 0, Z);
	printf("%s (kernel %s): ", name, sgemm_cuda_last_kernel() == UGEMM_MODE_3XTF32 ? "K1/3xTF32" : "K2/SIMT");
	double e = cmp_results(MSIZE, NSIZE, Z, C, NSIZE);
	printf("%s normwise relerr %.3e %s\n", name, e, e <= 1e-5 ? "ok" : "FAIL (> 1e-5)");
	return !(e <= 1e-5);
}

int main(void)
{
	if (sgemm_init(MSIZE * KSIZE, KSIZE * NSIZE, MSIZE * NSIZE)) { fprintf(stderr, "init: %s\n", sgemm_cuda_last_error()); return 1; }
	int bad = 0;
	{   /* known answer, sgemm_test.c:186-200 */
		float a[6] = {1, 2, 3, 4, 5, 6}, b[6] = {1, 2, 3, 4, 5, 6}, c[9] = {0}, want[9] = {9, 12, 15, 19, 26, 33, 29, 40, 51};
		sgemm_rnn(3, 3, 2, 1, a, b, 0, c);
		int ok = !memcmp(c, want, sizeof want);
		printf("3x2 . 2x3 known answer: %s\n", ok ? "ok" : "FAIL");
		bad |= !ok;
	}
	for (int i = 0; i < MSIZE * KSIZE; i++) A[i] = (float)(i + 1);
	for (int i = 0; i < KSIZE * NSIZE; i++) B[i] = (float)(i + 1);
	bad |= run_case("RNN", 0);
	bad |= run_case("RNT", 1);
	bad |= run_case("RTN", 2);
	sgemm_finish();
	return bad;
}
