/* check_sgemm_cuda.c -- the reference's check_sgemm.c harness extended to the CUDA backend.
 *
 * Keeps check_sgemm.c's shape: `key=value` argv grammar (check_sgemm.c:159-221: alpha= beta= M= N= K= lda= ldb=
 * ldc=), a list of `uut` function pointers with the shared 14-argument signature (check_sgemm.c:96-103), nIter
 * independent problem instances, a reference result, and the cmp_results line (check_sgemm.c:56-85) -- and fixes
 * what breaks at the BASELINE shapes: heap instead of a stack VLA for the reference result,
 * 64-bit sizes and flop counts, wall-clock timing instead of TSC x 3.5 GHz, seeded inputs, and a fast CPU reference
 * (sgemm_avx / sgemm_sse) instead of the naive loop above 256^3.  New keys: ta= tb= major= iters= seed= lo= hi=
 * mode=auto|3xtf32|simt|all  check=0|1.
 *
 * Rows printed: the reference CPU implementation (timed on this host, core count stated) and one row per CUDA entry
 * point (host-pointer call: H2D + kernel + D2H inside the timed region, like the reference's OpenCL numbers, plus the
 * device-resident kernel time).  Gate: normwise relative error <= 1e-5 (north_star); exit status 1 on failure.
 *
 * The CPU reference is TEST INFRASTRUCTURE loaded at run time with dlopen from oracle/_ref/libugemm_ref.so (the
 * unmodified reference compiled by oracle/Makefile) or, if that is absent, oracle/liboracle.so (our restatement).
 * The product library libugemm_cuda.so is linked normally and contains no CPU path.
 *
 * Build: make -C harness      Run: LD_LIBRARY_PATH=ugemm_b200 harness/check_sgemm_cuda M=1024 N=1024 K=1024
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "ugemm_cuda.h"

typedef void (*uut_t)(char major, char transa, char transb, int M, int N, int K, float alpha, const float *A, int lda,
                      const float *B, int ldb, float beta, float *C, int ldc);
typedef void (*mt_t)(int threads, char major, char transa, char transb, int M, int N, int K, float alpha, const float *A,
                     int lda, const float *B, int ldb, float beta, float *C, int ldc);

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* check_sgemm.c:56-85, with 64-bit indexing; also returns the normwise relative error. */
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
	printf("%.3e/%.3e=%.3e. %.3e at [%3zu,%3zu] %18.10e vs %18.10e %s\n", stdErr, stdRef, stdErr / stdRef, maxErr,
	       maxI / ld, maxI % ld, (double)ref[maxI], (double)res[maxI],
	       maxErr > stdRef * 1e-5 ? "FAIL !!!" : (maxErr > stdRef * 3e-5 || stdErr > stdRef * 1e-6 ? "Sucks !" : ""));
	return s2Ref > 0 ? sqrt(s2Err / s2Ref) : (s2Err > 0 ? INFINITY : 0);
}

static void *load_checker(const char **kind)
{
	const char *cands[] = {"oracle/_ref/libugemm_ref.so", "../oracle/_ref/libugemm_ref.so", "oracle/liboracle.so", "../oracle/liboracle.so"};
	for (unsigned i = 0; i < sizeof cands / sizeof *cands; i++) {
		void *h = dlopen(cands[i], RTLD_NOW | RTLD_LOCAL);
		if (h) { *kind = cands[i]; return h; }
	}
	return NULL;
}

int main(int argz, char **argv)
{
	float alpha = 1, beta = 0, lo = 0, hi = 1;
	int M = 128, N = 361, K = 1152; /* check_sgemm.c:147-154 defaults */
	int lda = 0, ldb = 0, ldc = 0, iters = 3, check = 1;
	unsigned long long seed = 1;
	char ta = 'N', tb = 'N', major = 'R';
	const char *mode = "all";
	for (int i = 1; i < argz; i++) {
		char *eq = strchr(argv[i], '=');
		if (!eq) { fprintf(stderr, "bad argument '%s' (want key=value)\n", argv[i]); return 2; }
		*eq = 0;
		const char *k = argv[i], *v = eq + 1;
		if (!strcmp(k, "alpha")) alpha = strtof(v, NULL);
		else if (!strcmp(k, "beta")) beta = strtof(v, NULL);
		else if (!strcmp(k, "M")) M = atoi(v);
		else if (!strcmp(k, "N")) N = atoi(v);
		else if (!strcmp(k, "K")) K = atoi(v);
		else if (!strcmp(k, "lda")) lda = atoi(v);
		else if (!strcmp(k, "ldb")) ldb = atoi(v);
		else if (!strcmp(k, "ldc")) ldc = atoi(v);
		else if (!strcmp(k, "ta")) ta = v[0];
		else if (!strcmp(k, "tb")) tb = v[0];
		else if (!strcmp(k, "major")) major = v[0];
		else if (!strcmp(k, "iters")) iters = atoi(v);
		else if (!strcmp(k, "seed")) seed = strtoull(v, NULL, 10);
		else if (!strcmp(k, "lo")) lo = strtof(v, NULL);
		else if (!strcmp(k, "hi")) hi = strtof(v, NULL);
		else if (!strcmp(k, "mode")) mode = v;
		else if (!strcmp(k, "check")) check = atoi(v);
		else { fprintf(stderr, "unknown key '%s'\n", k); return 2; }
	}
	const int rowmaj = major == 'R' || major == 'r';
	const int tA = ta == 'T' || ta == 't', tB = tb == 'T' || tb == 't';
	/* stored shapes: lines x width */
	const int a_lines = (rowmaj ? !tA : tA) ? M : K, a_w = (rowmaj ? !tA : tA) ? K : M;
	const int b_lines = (rowmaj ? !tB : tB) ? K : N, b_w = (rowmaj ? !tB : tB) ? N : K;
	const int c_lines = rowmaj ? M : N, c_w = rowmaj ? N : M;
	if (lda < a_w) lda = a_w; /* check_sgemm.c:223-225: tight by default */
	if (ldb < b_w) ldb = b_w;
	if (ldc < c_w) ldc = c_w;
	printf("major=%c ta=%c tb=%c M=%d N=%d K=%d alpha=%g beta=%g lda=%d ldb=%d ldc=%d seed=%llu U[%g,%g)\n", major, ta, tb, M, N, K,
	       alpha, beta, lda, ldb, ldc, seed, lo, hi);
	const double u = 5.9604644775390625e-08;
	printf("a-priori bounds: K*u = %.2e, sqrt(K)*u = %.2e; gate: normwise relerr <= 1e-5\n", K * u, sqrt((double)K) * u);

	if (sgemm_cuda_init(-1, 0)) { fprintf(stderr, "sgemm_cuda_init: %s\n", sgemm_cuda_last_error()); return 1; }
	int sms = 0, khz = 0; size_t hbm = 0; char name[128];
	ugemm_cuda_device_info(&sms, &khz, &hbm, name, sizeof name);
	printf("device: %s, %d SMs, %.0f MHz, %.0f GB\n", name, sms, khz / 1e3, hbm / 1e9);

	const size_t an = (size_t)a_lines * lda, bn = (size_t)b_lines * ldb, cn = (size_t)c_lines * ldc;
	float *A = ugemm_cuda_malloc_host(an * 4), *B = ugemm_cuda_malloc_host(bn * 4);
	float *C0 = malloc(cn * 4), *C = ugemm_cuda_malloc_host(cn * 4), *R = malloc(cn * 4);
	if (!A || !B || !C0 || !C || !R) { fprintf(stderr, "allocation failed\n"); return 1; }
	ugemm_fill_uniform_host(A, an, seed * 3 + 0, lo, hi);   /* random_matrix with a seed */
	ugemm_fill_uniform_host(B, bn, seed * 3 + 1, lo, hi);
	ugemm_fill_uniform_host(C0, cn, seed * 3 + 2, lo, hi);
	const double flop = 2.0 * M * N * K;
	int failed = 0;

	/* ---- CPU reference (checker): NN row-major -> sgemm_avx on all cores; anything else -> sgemm_sse (1 core) */
	int have_ref = 0;
	if (check) {
		const char *kind = NULL;
		void *h = load_checker(&kind);
		if (!h) { fprintf(stderr, "no checker library (oracle/_ref or oracle/liboracle.so): run `make -C oracle`\n"); return 1; }
		const int cores = (int)sysconf(_SC_NPROCESSORS_ONLN);
		mt_t avx_mt = (mt_t)dlsym(h, "ref_sgemm_avx_mt");
		uut_t sse = (uut_t)dlsym(h, "ref_sgemm_sse"), naive = (uut_t)dlsym(h, "ref_sgemm_cpu");
		mt_t banded = (mt_t)dlsym(h, "oracle_sgemm_banded");
		memcpy(R, C0, cn * 4);
		double t0 = now_s();
		const char *what;
		int used = 1;
		if ((size_t)M * N * K <= (size_t)256 * 256 * 256 && naive) { naive(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, R, ldc); what = "sgemm_cpu (naive, ugemm.h:287)"; }
		else if (avx_mt && rowmaj && !tA && !tB && alpha != 0) { avx_mt(cores, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, R, ldc); what = "sgemm_avx on row slabs (sgemm_avx256.h:392)"; used = cores; }
		else if (sse) { sse(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, R, ldc); what = "sgemm_sse (sgemm_sse.h:365)"; }
		else if (banded) { banded(cores, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, R, ldc); what = "oracle 35-band restatement"; used = cores; }
		else { fprintf(stderr, "checker library %s exports no usable SGEMM\n", kind); return 1; }
		double dt = now_s() - t0;
		printf("%-44s %10.3f ms %10.1f GFLOPS  (%d of %d host cores, %s)\n", what, dt * 1e3, flop / dt / 1e9, used, cores, kind);
		have_ref = 1;
	}

	struct { const char *name; uut_t fn; int mode; } rows[] = {
		{"sgemm_cuda_3xtf32", sgemm_cuda_3xtf32, UGEMM_MODE_3XTF32}, {"sgemm_cuda_simt", sgemm_cuda_simt, UGEMM_MODE_SIMT}, {"sgemm_cuda (auto)", sgemm_cuda, UGEMM_MODE_AUTO}};
	for (unsigned r = 0; r < 3; r++) {
		if (strcmp(mode, "all") && !((!strcmp(mode, "3xtf32") && r == 0) || (!strcmp(mode, "simt") && r == 1) || (!strcmp(mode, "auto") && r == 2))) continue;
		double best = 1e30;
		int err = 0;
		for (int it = 0; it < iters + 1 && !err; it++) {   /* first pass = warm-up */
			memcpy(C, C0, cn * 4);
			double t0 = now_s();
			rows[r].fn(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
			double dt = now_s() - t0;
			if (sgemm_cuda_last_error()) { printf("%-20s not run: %s\n", rows[r].name, sgemm_cuda_last_error()); sgemm_cuda_clear_error(); err = 1; }
			if (it && dt < best) best = dt;
		}
		if (err) continue;
		/* device-resident kernel time for the same problem */
		float *dA = ugemm_cuda_malloc(an * 4), *dB = ugemm_cuda_malloc(bn * 4), *dC = ugemm_cuda_malloc(cn * 4);
		float kavg = 0, kmin = 0;
		if (dA && dB && dC) {
			ugemm_cuda_memcpy_h2d(dA, A, an * 4); ugemm_cuda_memcpy_h2d(dB, B, bn * 4); ugemm_cuda_memcpy_h2d(dC, C0, cn * 4);
			sgemm_cuda_time_dev(rows[r].mode, iters > 0 ? iters : 1, 1, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta == 0 ? 0.f : beta, dC, ldc, &kavg, &kmin, NULL);
		}
		ugemm_cuda_free(dA); ugemm_cuda_free(dB); ugemm_cuda_free(dC);
		printf("%-20s kernel=%s  host-ptr call %9.3f ms %9.1f GFLOPS | device-resident %8.3f ms %9.1f GFLOPS\n", rows[r].name,
		       sgemm_cuda_last_kernel() == UGEMM_MODE_3XTF32 ? "K1/3xTF32" : "K2/SIMT", best * 1e3, flop / best / 1e9, kmin, kmin > 0 ? flop / kmin / 1e6 : 0.0);
		if (have_ref) {
			double worst = 0;
			if (rowmaj) worst = cmp_results(M, N, R, C, ldc);
			else worst = cmp_results(N, M, R, C, ldc);   /* column-major C is an N x M row-major array */
			printf("%-20s normwise relerr %.3e %s\n", rows[r].name, worst, worst <= 1e-5 ? "ok" : "FAIL (> 1e-5)");
			if (!(worst <= 1e-5)) failed = 1;
			/* ld padding must be untouched */
			for (int l = 0; l < c_lines && !failed; l++)
				for (int c = c_w; c < ldc; c++)
					if (C[(size_t)l * ldc + c] != C0[(size_t)l * ldc + c]) { printf("padding written at line %d col %d\n", l, c); failed = 1; break; }
		}
	}
	ugemm_cuda_free_host(A); ugemm_cuda_free_host(B); ugemm_cuda_free_host(C);
	free(C0); free(R);
	sgemm_cuda_finish();
	return failed;
}
