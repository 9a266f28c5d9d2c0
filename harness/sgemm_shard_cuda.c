/* sgemm_shard_cuda.c -- BASELINE config 5 from the C host side with ONE PROCESS PER GPU: the sharded SGEMM of the C ABI
 * (sgemm_cuda_shard_*, csrc/shard.cu) driven the way an MPI program would drive it, without MPI: the parent forks one child per
 * GPU before anything touches CUDA, and the only thing the children exchange outside the library is the 128-byte NCCL id, through
 * a shared page (what MPI_Bcast would carry).  In the shape of the reference's harnesses: generate, run, compare, print GFLOPS
 * (check_sgemm.c:87-143; flops = 2*M*N*K as check_sgemm.c:131).
 *
 *   ./sgemm_shard_cuda [nproc [M [N [K [steps [transport]]]]]]     default: every visible GPU, 32768^3, 5 steps, transport 1
 *
 * transport 1 = copy-engine peer pull over CUDA IPC, 0 = NCCL broadcast in grid-row / grid-column communicators.  Printed: the step
 * with panel distribution, the products alone (panels resident), the other transport, and the end-to-end step from pinned host
 * memory; every rank verifies sampled entries of its C block against double-precision dot products of regenerated operand windows
 * (normwise gate 1e-5).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "ugemm_cuda.h"

typedef struct {
	volatile int id_ready;
	unsigned char id[2][128];
	volatile int id2_ready;
	float with_ms, without_ms, other_ms, e2e_ms, floor_ms;
	int transport, other_transport, pr, pc, slabs;
	double relerr[16];
	int failed[16];
} Shared;

static void nap(void) { struct timespec ts = {0, 2000000}; nanosleep(&ts, NULL); }

static int rank_main(Shared *sh, int rank, int world, int M, int N, int K, int steps, int transport)
{
	if (sgemm_cuda_init(rank, 0)) { fprintf(stderr, "[rank %d] init failed: %s\n", rank, sgemm_cuda_last_error()); return 2; }
	if (rank == 0) {
		if (world > 1 && (sgemm_cuda_shard_unique_id(sh->id[0]) || sgemm_cuda_shard_unique_id(sh->id[1]))) { fprintf(stderr, "unique id: %s\n", sgemm_cuda_last_error()); return 2; }
		sh->id_ready = 1;
	}
	while (!sh->id_ready) nap();
	if (sgemm_cuda_shard_init(rank, world, world > 1 ? sh->id[0] : NULL, M, N, K, transport)) { fprintf(stderr, "[rank %d] shard init failed: %s\n", rank, sgemm_cuda_last_error()); return 2; }
	float with_ms = 0, without_ms = 0, e2e_ms = 0, floor_ms = 0, v;
	long long up = 0, down = 0;
	if (sgemm_cuda_shard_generate(1, 2, -0.5f, 0.5f) || sgemm_cuda_shard_run(1, steps, 3, &with_ms) || sgemm_cuda_shard_run(0, steps > 2 ? steps / 2 : 2, 1, &without_ms) ||
	    sgemm_cuda_shard_download_owned() || sgemm_cuda_shard_run_host(2, 1, &e2e_ms, &up, &down) || sgemm_cuda_shard_copy_floor(2, &floor_ms)) {
		fprintf(stderr, "[rank %d] run failed: %s\n", rank, sgemm_cuda_last_error());
		return 1;
	}
	v = with_ms / steps; sgemm_cuda_shard_allreduce(&v, 0); with_ms = v;
	v = without_ms / (steps > 2 ? steps / 2 : 2); sgemm_cuda_shard_allreduce(&v, 0); without_ms = v;
	v = e2e_ms / 2; sgemm_cuda_shard_allreduce(&v, 0); e2e_ms = v;
	v = floor_ms / 2; sgemm_cuda_shard_allreduce(&v, 0); floor_ms = v;

	/* sampled verification of this rank's block: 8 rows x 32 columns against fp64 dot products of regenerated windows */
	float *dC; int rows, cols, r0, c0;
	if (sgemm_cuda_shard_run(1, 1, 0, NULL) || sgemm_cuda_shard_block(&dC, &rows, &cols, &r0, &c0)) return 1;
	float *a = malloc((size_t)K * 4), *b = malloc((size_t)K * 4), c;
	double num = 0, den = 0;
	for (int t = 0; t < 32; t++) {
		const size_t col = ((size_t)t * cols) / 32 + (size_t)(t * 29) % (cols / 32 > 0 ? cols / 32 : 1);
		ugemm_fill_uniform_host_2d(b, (size_t)K, 1, 1, 2, c0 + col, (size_t)N, -0.5f, 0.5f);
		for (int s = 0; s < 8; s++) {
			const size_t row = ((size_t)s * rows) / 8 + (size_t)(s * 37 + t) % (rows / 8 > 0 ? rows / 8 : 1);
			ugemm_fill_uniform_host_2d(a, 1, (size_t)K, (size_t)K, 1, (r0 + row) * (size_t)K, (size_t)K, -0.5f, 0.5f);
			ugemm_cuda_memcpy_d2h(&c, dC + row * (size_t)cols + col, sizeof(float));
			double acc = 0;
			for (int k = 0; k < K; k++) acc += (double)a[k] * (double)b[k];
			num += (c - acc) * (c - acc);
			den += acc * acc;
		}
	}
	sh->relerr[rank] = sqrt(num / den);
	sh->failed[rank] = !(sh->relerr[rank] <= 1e-5);
	const int used = sgemm_cuda_shard_transport();
	sgemm_cuda_shard_finish();

	/* the other transport on the same problem */
	float other_ms = 0;
	const int other = used == 1 ? 0 : 1;
	int other_used = -1;
	if (world > 1) {
		if (sgemm_cuda_shard_init(rank, world, sh->id[1], M, N, K, other)) { fprintf(stderr, "[rank %d] second init failed: %s\n", rank, sgemm_cuda_last_error()); return 2; }
		other_used = sgemm_cuda_shard_transport();
		if (sgemm_cuda_shard_generate(1, 2, -0.5f, 0.5f) || sgemm_cuda_shard_run(1, steps, 2, &other_ms)) return 1;
		v = other_ms / steps; sgemm_cuda_shard_allreduce(&v, 0); other_ms = v;
		sgemm_cuda_shard_finish();
	}
	if (rank == 0) {
		sh->with_ms = with_ms; sh->without_ms = without_ms; sh->e2e_ms = e2e_ms; sh->floor_ms = floor_ms; sh->other_ms = other_ms;
		sh->transport = used; sh->other_transport = other_used;
		sgemm_cuda_shard_plan(world, 0, M, N, K, &sh->pr, &sh->pc, &sh->slabs, NULL, NULL, NULL);
	}
	free(a); free(b);
	sgemm_cuda_finish();
	return sh->failed[rank] ? 1 : 0;
}

int main(int argc, char **argv)
{
	int world = argc > 1 ? atoi(argv[1]) : 0;
	const int M = argc > 2 ? atoi(argv[2]) : 32768;
	const int N = argc > 3 ? atoi(argv[3]) : M;
	const int K = argc > 4 ? atoi(argv[4]) : M;
	const int steps = argc > 5 && atoi(argv[5]) > 0 ? atoi(argv[5]) : 5;
	const int transport = argc > 6 ? atoi(argv[6]) : 1;
	if (world <= 0) {
		/* the device count is asked in a child so that the parent never initialises CUDA before it forks */
		int fd[2];
		if (pipe(fd)) return 2;
		pid_t p = fork();
		if (p == 0) { int n = ugemm_cuda_device_count(); if (write(fd[1], &n, sizeof n) != sizeof n) _exit(2); _exit(0); }
		if (read(fd[0], &world, sizeof world) != sizeof world) world = 0;
		waitpid(p, NULL, 0);
		world = world >= 8 ? 8 : world >= 4 ? 4 : world >= 2 ? 2 : world;
	}
	if (world < 1 || world > 16) { fprintf(stderr, "no CUDA device visible\n"); return 2; }
	Shared *sh = mmap(NULL, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	if (sh == MAP_FAILED) { perror("mmap"); return 2; }
	memset(sh, 0, sizeof *sh);
	pid_t pids[16];
	for (int r = 0; r < world; r++) {
		pids[r] = fork();
		if (pids[r] < 0) { perror("fork"); return 2; }
		if (pids[r] == 0) _exit(rank_main(sh, r, world, M, N, K, steps, transport));
	}
	int bad = 0;
	for (int r = 0; r < world; r++) {
		int st = 0;
		waitpid(pids[r], &st, 0);
		if (!WIFEXITED(st) || WEXITSTATUS(st)) bad = 1;
	}
	if (bad) { fprintf(stderr, "a rank failed\n"); return 1; }
	const double flops = 2.0 * M * N * (double)K;
	const char *names[2] = {"NCCL broadcast", "copy-engine peer pull"};
	printf("sgemm_cuda_shard  M=%d N=%d K=%d  grid %d x %d (%d processes, one GPU each)  %d K slabs  row-major NN alpha=1 beta=0\n", M, N, K, sh->pr, sh->pc, world, sh->slabs);
	printf(">>> with panel distribution (%s): %.3f ms  %.1f TFLOP/s   (device time, max over ranks)\n", names[sh->transport], sh->with_ms, flops / sh->with_ms / 1e9);
	printf(">>> products only (panels resident):          %.3f ms  %.1f TFLOP/s\n", sh->without_ms, flops / sh->without_ms / 1e9);
	if (world > 1 && sh->other_transport >= 0)
		printf(">>> with panel distribution (%s): %.3f ms  %.1f TFLOP/s\n", names[sh->other_transport], sh->other_ms, flops / sh->other_ms / 1e9);
	printf(">>> end to end from pinned host memory:       %.3f ms  %.1f TFLOP/s   (host link floor of the same bytes: %.3f ms)\n", sh->e2e_ms, flops / sh->e2e_ms / 1e9, sh->floor_ms);
	double worst = 0;
	for (int r = 0; r < world; r++) if (sh->relerr[r] > worst) worst = sh->relerr[r];
	printf("sampled relerr, worst rank (8 x 32 entries per rank): %.3e (gate 1e-5)  %s\n", worst, worst <= 1e-5 ? "ok" : "FAIL !!!");
	return worst <= 1e-5 ? 0 : 1;
}
