/* check_sgemm_cuda.c -- the reference's check_sgemm.c harness extended to the CUDA backend.
 *
 * Keeps check_sgemm.c's shape: `key=value` argv grammar (check_sgemm.c:159-221: alpha= beta= M= N= K= lda= ldb=
 * ldc=), a list of `uut` function pointers with the shared 14-argument signature (check_sgemm.c:96-103), `inst`
 * independent problem instances STACKED in one buffer each and walked one call at a time (test_sgemm,
 * check_sgemm.c:111-124; the reference's nIter = 11 is `inst=11`), a reference result per instance, and the
 * cmp_results line (check_sgemm.c:56-85) -- and fixes what breaks at the BASELINE shapes: heap
 * instead of a stack VLA for the reference result, 64-bit sizes and flop counts, wall-clock timing instead of
 * TSC x 3.5 GHz, seeded inputs, and a fast CPU reference (sgemm_avx / sgemm_sse) instead of the naive loop above 256^3.
 * New keys: ta= tb= major= inst= iters= seed= lo= hi= check=0|1 and
 *     mode=all|auto|3xtf32|simt   the CUDA rows to run (default all); a missing GPU is an ERROR (exit 1), never a fallback
 *     mode=cpu                    host-only pass, no GPU touched: the reference's own uut rows (check_sgemm.c:247-250:
 *                                 sgemm_c, sgemm_avx as shipped on ONE core, sgemm_sse) plus sgemm_avx on row slabs over
 *                                 all cores, each compared with the ground truth; the CUDA rows are reported "not run".
 *                                 This is BASELINE config 1 (`mode=cpu M=1024 N=1024 K=1024`).
 *
 * Rows printed in the CUDA modes: the reference CPU implementation (timed on this host, core count stated) and one row
 * per CUDA entry point (host-pointer call per instance: H2D + kernel + D2H inside the timed region, like the reference's
 * OpenCL numbers, plus the device-resident kernel time), and for inst > 1 one more row for the whole stack in ONE call
 * (sgemm_cuda_batched).  Gate: normwise relative error <= 1e-5 on every instance (north_star); exit status 1 on failure.
 *
 * The CPU reference is TEST INFRASTRUCTURE loaded at run time with dlopen from oracle/_ref/libugemm_ref.so (the
 * unmodified reference compiled by oracle/Makefile) or, if that is absent, oracle/liboracle.so (our restatement).
 * The product library libugemm_cuda.so is linked normally and contains no CPU path.
 *
 * Build: make -C harness      Run: harness/check_sgemm_cuda M=1024 N=1024 K=1024 inst=11
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

/* problem description shared by the row runners */
static struct {
	char major, ta, tb;
	int M, N, K, lda, ldb, ldc, inst, c_lines, c_w, rowmaj;
	float alpha, beta;
	size_t an, bn, cn;          /* floats per instance of A, B, C */
	const float *A, *B, *C0;    /* stacked inputs */
	float *C, *R;               /* stacked result / reference result */
	int have_ref;
	double flop;
} P;

/* test_sgemm (check_sgemm.c:87-143): restore C from the source copy, walk the stacked instances one call at a time (each
 * timed), print the per-instance times and their mean, then compare every instance with the reference result.
 * Returns the worst normwise relative error (or -1 when an instance reported an error through `failed_call`). */
static double walk_instances(const char *name, uut_t fn, mt_t fn_mt, int threads, int (*failed_call)(void), double *mean_s)
{
	memcpy(P.C, P.C0, P.cn * P.inst * 4);
	double sum = 0;
	printf("%-28s", name);
	for (int it = 0; it < P.inst; it++) {
		double t0 = now_s();
		if (fn_mt) fn_mt(threads, P.major, P.ta, P.tb, P.M, P.N, P.K, P.alpha, P.A + it * P.an, P.lda, P.B + it * P.bn, P.ldb, P.beta, P.C + it * P.cn, P.ldc);
		else fn(P.major, P.ta, P.tb, P.M, P.N, P.K, P.alpha, P.A + it * P.an, P.lda, P.B + it * P.bn, P.ldb, P.beta, P.C + it * P.cn, P.ldc);
		double dt = now_s() - t0;
		if (failed_call && failed_call()) { printf(" not run\n"); return -1; }
		if (P.inst <= 11) printf(" %.3f", dt * 1e3);
		sum += dt;
	}
	*mean_s = sum / P.inst;
	printf("%s: mean %.3f ms %.1f GFLOPS\n", P.inst <= 11 ? " ms" : "", *mean_s * 1e3, P.flop / *mean_s / 1e9);
	return 0;
}

static double compare_instances(const char *name)
{
	double worst = 0;
	for (int it = 0; it < P.inst; it++) {
		double e = P.rowmaj ? cmp_results(P.M, P.N, P.R + it * P.cn, P.C + it * P.cn, P.ldc)
		                    : cmp_results(P.N, P.M, P.R + it * P.cn, P.C + it * P.cn, P.ldc);   /* column-major C is an N x M row-major array */
		if (!(e <= worst)) worst = e;
	}
	printf("%-28s worst normwise relerr over %d instance%s %.3e %s\n", name, P.inst, P.inst > 1 ? "s" : "", worst, worst <= 1e-5 ? "ok" : "FAIL (> 1e-5)");
	return worst;
}

static int padding_untouched(void)
{
	for (int it = 0; it < P.inst; it++)
		for (int l = 0; l < P.c_lines; l++)
			for (int c = P.c_w; c < P.ldc; c++)
				if (P.C[it * P.cn + (size_t)l * P.ldc + c] != P.C0[it * P.cn + (size_t)l * P.ldc + c]) {
					printf("padding written at instance %d line %d col %d\n", it, l, c);
					return 0;
				}
	return 1;
}

static int cuda_call_failed(void)
{
	if (!sgemm_cuda_last_error()) return 0;
	printf(" -- %s;", sgemm_cuda_last_error());
	sgemm_cuda_clear_error();
	return 1;
}

int main(int argz, char **argv)
{
	float alpha = 1, beta = 0, lo = 0, hi = 1;
	int M = 128, N = 361, K = 1152; /* check_sgemm.c:147-154 defaults */
	int lda = 0, ldb = 0, ldc = 0, iters = 3, check = 1, inst = 1;
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
		else if (!strcmp(k, "inst")) inst = atoi(v);
		else if (!strcmp(k, "iters")) iters = atoi(v);
		else if (!strcmp(k, "seed")) seed = strtoull(v, NULL, 10);
		else if (!strcmp(k, "lo")) lo = strtof(v, NULL);
		else if (!strcmp(k, "hi")) hi = strtof(v, NULL);
		else if (!strcmp(k, "mode")) mode = v;
		else if (!strcmp(k, "check")) check = atoi(v);
		else { fprintf(stderr, "unknown key '%s'\n", k); return 2; }
	}
	if (inst < 1 || inst > 64 || M < 1 || N < 1 || K < 1) { fprintf(stderr, "need 1 <= inst <= 64 and positive M, N, K\n"); return 2; }
	const int cpu_only = !strcmp(mode, "cpu");
	const int rowmaj = major == 'R' || major == 'r';
	const int tA = ta == 'T' || ta == 't', tB = tb == 'T' || tb == 't';
	/* stored shapes: lines x width */
	const int a_lines = (rowmaj ? !tA : tA) ? M : K, a_w = (rowmaj ? !tA : tA) ? K : M;
	const int b_lines = (rowmaj ? !tB : tB) ? K : N, b_w = (rowmaj ? !tB : tB) ? N : K;
	const int c_lines = rowmaj ? M : N, c_w = rowmaj ? N : M;
	if (lda < a_w) lda = a_w; /* check_sgemm.c:223-225: tight by default */
	if (ldb < b_w) ldb = b_w;
	if (ldc < c_w) ldc = c_w;
	printf("major=%c ta=%c tb=%c M=%d N=%d K=%d alpha=%g beta=%g lda=%d ldb=%d ldc=%d inst=%d seed=%llu U[%g,%g)\n", major, ta, tb, M, N, K,
	       alpha, beta, lda, ldb, ldc, inst, seed, lo, hi);
	const double u = 5.9604644775390625e-08;
	printf("a-priori bounds: K*u = %.2e, sqrt(K)*u = %.2e; gate: normwise relerr <= 1e-5\n", K * u, sqrt((double)K) * u);

	if (!cpu_only) {
		if (sgemm_cuda_init(-1, 0)) { fprintf(stderr, "sgemm_cuda_init: %s (mode=cpu runs the host-only reference pass)\n", sgemm_cuda_last_error()); return 1; }
		int sms = 0, khz = 0; size_t hbm = 0; char name[128];
		ugemm_cuda_device_info(&sms, &khz, &hbm, name, sizeof name);
		printf("device: %s, %d SMs, %.0f MHz, %.0f GB\n", name, sms, khz / 1e3, hbm / 1e9);
	}

	/* stacked instances (check_sgemm.c:242-246) */
	const size_t an = (size_t)a_lines * lda, bn = (size_t)b_lines * ldb, cn = (size_t)c_lines * ldc;
	float *A, *B, *C;
	if (cpu_only) { A = aligned_alloc(64, (an * inst * 4 + 63) / 64 * 64); B = aligned_alloc(64, (bn * inst * 4 + 63) / 64 * 64); C = aligned_alloc(64, (cn * inst * 4 + 63) / 64 * 64); }
	else { A = ugemm_cuda_malloc_host(an * inst * 4); B = ugemm_cuda_malloc_host(bn * inst * 4); C = ugemm_cuda_malloc_host(cn * inst * 4); }
	float *C0 = malloc(cn * inst * 4), *R = malloc(cn * inst * 4);
	if (!A || !B || !C0 || !C || !R) { fprintf(stderr, "allocation failed\n"); return 1; }
	ugemm_fill_uniform_host(A, an * inst, seed * 3 + 0, lo, hi);   /* random_matrix with a seed */
	ugemm_fill_uniform_host(B, bn * inst, seed * 3 + 1, lo, hi);
	ugemm_fill_uniform_host(C0, cn * inst, seed * 3 + 2, lo, hi);
	P.major = major; P.ta = ta; P.tb = tb; P.M = M; P.N = N; P.K = K; P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.inst = inst;
	P.c_lines = c_lines; P.c_w = c_w; P.rowmaj = rowmaj; P.alpha = alpha; P.beta = beta; P.an = an; P.bn = bn; P.cn = cn;
	P.A = A; P.B = B; P.C0 = C0; P.C = C; P.R = R; P.flop = 2.0 * M * N * K;
	int failed = 0;
	const int cores = (int)sysconf(_SC_NPROCESSORS_ONLN);
	const int nn_row = rowmaj && !tA && !tB;

	/* ---- CPU reference (checker) ---- */
	void *h = NULL;
	const char *kind = NULL;
	uut_t sse = NULL, naive = NULL, goto_c = NULL, avx = NULL;
	mt_t avx_mt = NULL, banded = NULL;
	if (check || cpu_only) {
		h = load_checker(&kind);
		if (!h) { fprintf(stderr, "no checker library (oracle/_ref or oracle/liboracle.so): run `make -C oracle`\n"); return 1; }
		avx_mt = (mt_t)dlsym(h, "ref_sgemm_avx_mt");
		sse = (uut_t)dlsym(h, "ref_sgemm_sse"); naive = (uut_t)dlsym(h, "ref_sgemm_cpu");
		goto_c = (uut_t)dlsym(h, "ref_sgemm_c"); avx = (uut_t)dlsym(h, "ref_sgemm_avx");
		banded = (mt_t)dlsym(h, "oracle_sgemm_banded");
		/* ground truth for this run: the naive loop while it is affordable (always, in the reference: check_sgemm.c:137-141),
		 * in the host-only pass up to 2^31 multiply-adds; then sgemm_avx on row slabs (row-major NN) or sgemm_sse */
		const double work = (double)M * N * K * inst;
		const char *what;
		double mean = 0;
		uut_t tf = NULL; mt_t tm = NULL; int used = 1;
		if (naive && work <= (cpu_only ? 2147483648.0 : 16777216.0)) { tf = naive; what = "sgemm_cpu (naive, ugemm.h:287)"; }
		else if (avx_mt && nn_row && alpha != 0 && !cpu_only) { tm = avx_mt; used = cores; what = "sgemm_avx row slabs (sgemm_avx256.h:392)"; }
		else if (sse) { tf = sse; what = "sgemm_sse (sgemm_sse.h:365)"; }
		else if (banded) { tm = banded; used = cores; what = "oracle 35-band restatement"; }
		else { fprintf(stderr, "checker library %s exports no usable SGEMM\n", kind); return 1; }
		printf("reference result: %s, %d of %d host cores, %s\n", what, used, cores, kind);
		walk_instances("  reference", tf, tm, used, NULL, &mean);
		memcpy(R, C, cn * inst * 4);
		P.have_ref = 1;
	}

	if (cpu_only) {
		/* BASELINE config 1: the reference's own rows (check_sgemm.c:247-250), timed on this host, no GPU */
		struct { const char *name; uut_t fn; mt_t mt; int threads; int ok; } rows[] = {
			{"sgemm_c (gemm_cpu.h:284)", goto_c, NULL, 1, goto_c != NULL},
			{"sgemm_avx 1 core (as shipped)", avx, NULL, 1, avx != NULL && nn_row},     /* row-major NN only (sgemm_avx256.h:434) */
			{"sgemm_sse (sgemm_sse.h:365)", sse, NULL, 1, sse != NULL},
			{"sgemm_avx row slabs, all cores", NULL, avx_mt, cores, avx_mt != NULL && nn_row},
		};
		for (unsigned r = 0; r < sizeof rows / sizeof *rows; r++) {
			if (!rows[r].ok) { printf("%-28s not applicable to this case / not exported by %s\n", rows[r].name, kind); continue; }
			double mean = 0;
			walk_instances(rows[r].name, rows[r].fn, rows[r].mt, rows[r].threads, NULL, &mean);
			if (!(compare_instances(rows[r].name) <= 1e-5) || !padding_untouched()) failed = 1;
		}
		printf("sgemm_cuda_3xtf32 / sgemm_cuda_simt / sgemm_cuda: not run (mode=cpu; no GPU is touched in this pass and there is no CPU fallback in the CUDA library)\n");
		free(A); free(B); free(C); free(C0); free(R);
		return failed;
	}

	struct { const char *name; uut_t fn; int mode; } rows[] = {
		{"sgemm_cuda_3xtf32", sgemm_cuda_3xtf32, UGEMM_MODE_3XTF32}, {"sgemm_cuda_simt", sgemm_cuda_simt, UGEMM_MODE_SIMT}, {"sgemm_cuda (auto)", sgemm_cuda, UGEMM_MODE_AUTO}};
	for (unsigned r = 0; r < 3; r++) {
		if (strcmp(mode, "all") && !((!strcmp(mode, "3xtf32") && r == 0) || (!strcmp(mode, "simt") && r == 1) || (!strcmp(mode, "auto") && r == 2))) continue;
		double best = 1e30, mean = 0;
		int err = 0;
		for (int it = 0; it < iters + 1 && !err; it++) {   /* first pass = warm-up */
			if (walk_instances(rows[r].name, rows[r].fn, NULL, 1, cuda_call_failed, &mean) < 0) err = 1;
			if (it && mean < best) best = mean;
		}
		if (err) continue;
		/* device-resident kernel time for the same problem (first instance) */
		float *dA = ugemm_cuda_malloc(an * 4), *dB = ugemm_cuda_malloc(bn * 4), *dC = ugemm_cuda_malloc(cn * 4);
		float kavg = 0, kmin = 0;
		if (dA && dB && dC) {
			ugemm_cuda_memcpy_h2d(dA, A, an * 4); ugemm_cuda_memcpy_h2d(dB, B, bn * 4); ugemm_cuda_memcpy_h2d(dC, C0, cn * 4);
			sgemm_cuda_time_dev(rows[r].mode, iters > 0 ? iters : 1, 1, major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta == 0 ? 0.f : beta, dC, ldc, &kavg, &kmin, NULL);
		}
		ugemm_cuda_free(dA); ugemm_cuda_free(dB); ugemm_cuda_free(dC);
		printf("%-28s kernel=%s  host-ptr call %9.3f ms %9.1f GFLOPS | device-resident %8.3f ms %9.1f GFLOPS\n", rows[r].name,
		       sgemm_cuda_last_kernel() == UGEMM_MODE_3XTF32 ? "K1/3xTF32" : "K2/SIMT", best * 1e3, P.flop / best / 1e9, kmin, kmin > 0 ? P.flop / kmin / 1e6 : 0.0);
		if (P.have_ref && (!(compare_instances(rows[r].name) <= 1e-5) || !padding_untouched())) failed = 1;
	}
	if (inst > 1 && (!strcmp(mode, "all") || !strcmp(mode, "auto"))) {
		/* the whole stack in ONE call: sgemm_cuda_batched replaces test_sgemm's loop (check_sgemm.c:111-124) */
		double best = 1e30;
		int err = 0;
		for (int it = 0; it < iters + 1 && !err; it++) {
			memcpy(C, C0, cn * inst * 4);
			double t0 = now_s();
			sgemm_cuda_batched(major, ta, tb, M, N, K, alpha, A, lda, (long long)an, B, ldb, (long long)bn, beta, C, ldc, (long long)cn, inst);
			double dt = (now_s() - t0) / inst;
			if (sgemm_cuda_last_error()) { printf("%-28s not run: %s\n", "sgemm_cuda_batched", sgemm_cuda_last_error()); sgemm_cuda_clear_error(); err = 1; }
			if (it && dt < best) best = dt;
		}
		if (!err) {
			printf("%-28s kernel=%s  one call for %d instances: %9.3f ms per instance %9.1f GFLOPS\n", "sgemm_cuda_batched",
			       sgemm_cuda_last_kernel() == UGEMM_MODE_3XTF32 ? "K1/3xTF32" : "K2/SIMT", inst, best * 1e3, P.flop / best / 1e9);
			if (P.have_ref && (!(compare_instances("sgemm_cuda_batched") <= 1e-5) || !padding_untouched())) failed = 1;
		}
	}
	ugemm_cuda_free_host(A); ugemm_cuda_free_host(B); ugemm_cuda_free_host(C);
	free(C0); free(R);
	sgemm_cuda_finish();
	return failed;
}
