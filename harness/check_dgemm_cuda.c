/* check_dgemm_cuda.c -- the reference's check_dgemm.c harness with the CUDA backend as one more `uut` row.
 *
 * Keeps check_dgemm.c's shape: `key=value` argv grammar (alpha= beta= M= N= K= lda= ldb= ldc=), nIter stacked problem
 * instances walked one call at a time through a function pointer with the 14-argument double signature
 * (check_dgemm.c:86-125), a reference result per instance and the cmp_results line (check_dgemm.c:56-84) -- with the
 * fixes the CUDA-sized shapes need: heap instead of the stack VLA (check_dgemm.c:133), wall clock instead of TSC, seeded
 * inputs, and the reference's own blocked dgemm_avx instead of the naive dgemm_cpu as the CPU row above 512^3.
 * New keys: iters= seed= check=0|1.  Gate: normwise relative error <= 2e-14 (168 units of fp64 round-off, the analogue
 * of the SGEMM path's 1e-5); exit status 1 on failure.
 *
 * The CPU rows are TEST INFRASTRUCTURE loaded with dlopen from oracle/_ref/libugemm_ref.so (the unmodified reference) or
 * oracle/liboracle.so (our restatement); libugemm_cuda.so is linked normally and contains no CPU path.
 *
 * Build: make -C harness      Run: harness/check_dgemm_cuda M=1024 N=1024 K=1024
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ugemm_cuda.h"

typedef void (*duut_t)(char major, char transa, char transb, int M, int N, int K, double alpha, const double *A, int lda,
                       const double *B, int ldb, double beta, double *C, int ldc);

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* check_dgemm.c:56-84 with 64-bit indexing; returns the normwise relative error */
static double cmp_results(int M, int N, const double *ref, const double *res, int ld)
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
	printf("%.3e/%.3e=%.3e. %.3e at [%3zu,%3zu] %18.10e vs %18.10e\n", stdErr, stdRef, stdErr / stdRef, maxErr, maxI / ld, maxI % ld,
	       ref[maxI], res[maxI]);
	return s2Ref > 0 ? sqrt(s2Err / s2Ref) : (s2Err > 0 ? INFINITY : 0);
}

/* counter-based doubles in [0,1) with all 53 mantissa bits populated (splitmix64, like ugemm_fill_uniform_host) */
static void fill(double *x, size_t n, unsigned long long seed)
{
	unsigned long long s = seed * 0x9E3779B97F4A7C15ull;
	for (size_t i = 0; i < n; i++) {
		unsigned long long z = s + (i + 1) * 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z ^= z >> 31;
		x[i] = (double)(z >> 11) * 0x1p-53;
	}
}

static void run_rows(const char *name, duut_t f, int nIter, int M, int N, int K, double alpha, const double *a, int lda, const double *b,
                     int ldb, double beta, double *c, const double *sc, int ldc, int iters)
{
	memcpy(c, sc, sizeof(double) * (size_t)nIter * M * ldc);
	double best = 1e30;
	for (int r = 0; r < iters; r++)
		for (int it = 0; it < nIter; it++) {   /* check_dgemm.c:108-123 */
			if (r) memcpy(c + (size_t)it * M * ldc, sc + (size_t)it * M * ldc, sizeof(double) * (size_t)M * ldc);
			double t0 = now_s();
			f('R', 'N', 'N', M, N, K, alpha, a + (size_t)it * M * lda, lda, b + (size_t)it * K * ldb, ldb, beta, c + (size_t)it * M * ldc, ldc);
			double dt = now_s() - t0;
			if (dt < best) best = dt;
		}
	printf("%-34s best %.3f ms  %.2f GFLOP/s\n", name, best * 1e3, 2.0 * M * N * K / best / 1e9);
}

int main(int argz, char **argv)
{
	double alpha = 1, beta = 0;
	int M = 128, N = 361, K = 1152, lda = 0, ldb = 0, ldc = 0, iters = 2, check = 1;
	unsigned long long seed = 1;
	for (int i = 1; i < argz; i++) {
		char *eq = strchr(argv[i], '=');
		if (!eq) { fprintf(stderr, "bad argument '%s' (want key=value)\n", argv[i]); return 2; }
		*eq = 0;
		const char *k = argv[i], *v = eq + 1;
		if (!strcmp(k, "alpha")) alpha = strtod(v, NULL);
		else if (!strcmp(k, "beta")) beta = strtod(v, NULL);
		else if (!strcmp(k, "M")) M = atoi(v);
		else if (!strcmp(k, "N")) N = atoi(v);
		else if (!strcmp(k, "K")) K = atoi(v);
		else if (!strcmp(k, "lda")) lda = atoi(v);
		else if (!strcmp(k, "ldb")) ldb = atoi(v);
		else if (!strcmp(k, "ldc")) ldc = atoi(v);
		else if (!strcmp(k, "iters")) iters = atoi(v);
		else if (!strcmp(k, "seed")) seed = strtoull(v, NULL, 10);
		else if (!strcmp(k, "check")) check = atoi(v);
		else { fprintf(stderr, "unknown key '%s'\n", k); return 2; }
	}
	if (lda < K) lda = K;
	if (ldb < N) ldb = N;
	if (ldc < N) ldc = N;
	printf("Running DGEMM with M=%d, N=%d, K=%d, alpha=%f, lda=%d, ldb=%d, beta=%f, ldc=%d\n", M, N, K, alpha, lda, ldb, beta, ldc);
	printf("a-priori bounds: K*u = %.2e, sqrt(K)*u = %.2e (u = 2^-53); gate: normwise relerr <= 2e-14\n", K * 0x1p-53, sqrt((double)K) * 0x1p-53);
	if (sgemm_cuda_init(-1, 0)) { fprintf(stderr, "sgemm_cuda_init: %s\n", sgemm_cuda_last_error()); return 1; }

	const int nIter = 3;   /* check_dgemm.c uses 11; three keep the large shapes in host memory */
	double *a = malloc(sizeof(double) * (size_t)nIter * M * lda), *b = malloc(sizeof(double) * (size_t)nIter * K * ldb);
	double *c = malloc(sizeof(double) * (size_t)nIter * M * ldc), *sc = malloc(sizeof(double) * (size_t)nIter * M * ldc);
	double *ref = malloc(sizeof(double) * (size_t)nIter * M * ldc);
	if (!a || !b || !c || !sc || !ref) { fprintf(stderr, "out of host memory\n"); return 1; }
	fill(a, (size_t)nIter * M * lda, seed); fill(b, (size_t)nIter * K * ldb, seed + 1); fill(sc, (size_t)nIter * M * ldc, seed + 2);

	/* CPU row: the reference's dgemm_avx (or naive dgemm_cpu / our restatement), also the ground truth for the gate */
	duut_t cpu = NULL;
	const char *cpu_name = NULL;
	const char *cands[] = {"oracle/_ref/libugemm_ref.so", "../oracle/_ref/libugemm_ref.so", "oracle/liboracle.so", "../oracle/liboracle.so"};
	for (unsigned i = 0; i < 4 && !cpu; i++) {
		void *h = dlopen(cands[i], RTLD_NOW | RTLD_LOCAL);
		if (!h) continue;
		if ((cpu = (duut_t)dlsym(h, "ref_dgemm_avx"))) cpu_name = "dgemm_avx (reference, 1 core)";
		else if ((cpu = (duut_t)dlsym(h, "oracle_dgemm_naive"))) cpu_name = "oracle_dgemm_naive (restatement)";
	}
	int fail = 0;
	if (cpu && check) {
		run_rows(cpu_name, cpu, nIter, M, N, K, alpha, a, lda, b, ldb, beta, ref, sc, ldc, 1);
	} else if (check) {
		fprintf(stderr, "no CPU checker library found (oracle/_ref or oracle/liboracle.so): run `make -C oracle`\n");
		check = 0;
	}
	run_rows("dgemm_cuda (host ptrs, H2D+K4+D2H)", dgemm_cuda, nIter, M, N, K, alpha, a, lda, b, ldb, beta, c, sc, ldc, iters);
	if (sgemm_cuda_last_error()) { fprintf(stderr, "ugemm_cuda: %s\n", sgemm_cuda_last_error()); return 1; }
	for (int it = 0; it < nIter && check; it++) {
		double e = cmp_results(M, N, ref + (size_t)it * M * ldc, c + (size_t)it * M * ldc, ldc);
		printf("  instance %d: normwise relerr %.3e %s\n", it, e, e <= 2e-14 ? "ok" : "FAIL");
		if (!(e <= 2e-14)) fail = 1;
	}
	/* device-resident kernel time (what the host-pointer row hides behind PCIe) */
	double *dA = ugemm_cuda_malloc(sizeof(double) * (size_t)M * lda), *dB = ugemm_cuda_malloc(sizeof(double) * (size_t)K * ldb);
	double *dC = ugemm_cuda_malloc(sizeof(double) * (size_t)M * ldc);
	if (dA && dB && dC) {
		float avg = 0, best = 0;
		ugemm_cuda_memcpy_h2d(dA, a, sizeof(double) * (size_t)M * lda);
		ugemm_cuda_memcpy_h2d(dB, b, sizeof(double) * (size_t)K * ldb);
		if (!dgemm_cuda_time_dev(10, 3, 'R', 'N', 'N', M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &avg, &best))
			printf("%-34s avg %.3f ms best %.3f ms  %.2f TFLOP/s (FP64 pipe peak 148 SMs x 64 x 2 x 1.965 GHz = 37.2)\n", "dgemm_cuda_dev (device resident)",
			       avg, best, 2.0 * M * N * K / avg / 1e9);
	}
	ugemm_cuda_free(dA); ugemm_cuda_free(dB); ugemm_cuda_free(dC);
	sgemm_cuda_finish();
	free(a); free(b); free(c); free(sc); free(ref);
	printf(fail ? "FAILED\n" : "PASSED\n");
	return fail;
}
