// backend.cu -- the extern "C" boundary of the CUDA backend (include/ugemm_cuda.h).
//
// Replaces the reference's device-setup layers (ocl.h:141-360 oclSetup/oclKernel/oclKernelArgs/oclWrite/
// oclRead/oclRun/oclFinish; gpgpu_gl4.h:42-174 coInit/coCreateBuffer/coRun/coRead/coWrite/coTerm) with a thin
// CUDA runtime layer, and the per-call body of sgemm_ocl (sgemm_ocl2.h:166-218) with: validate -> normalise
// to row-major -> rule-based kernel choice (K1 3xTF32 tcgen05 | K2 SIMT FFMA) -> launch.  There is no CPU
// fallback anywhere in this file: if the device, the driver entry point or a launch fails, the error is
// recorded (sgemm_cuda_last_error) and the call returns without touching C.
#include "../../include/ugemm_cuda.h"
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>

using namespace ugemm;

namespace {

struct State {
	std::atomic<bool> ready{false};
	int device = 0;
	int sm_count = 0;
	int clock_khz = 0;
	size_t hbm_bytes = 0;
	char name[128] = {0};
	cudaStream_t stream = nullptr;   // compute + H2D
	cudaStream_t stream_d2h = nullptr;
	cudaStream_t stream_h2d = nullptr;
	cudaEvent_t ev_up[16] = {0}, ev_done[16] = {0};
	// staging arena of the host-pointer entry points (one buffer like the OpenCL backend's `gm`)
	char *arena = nullptr;
	size_t arena_bytes = 0;
	K1Tuning tuning = {4, 0, 0, 1};   // promote every 4 k-blocks (128 k), truncation split, CTA pairing by problem size, A collector
	// reports: written by whichever host thread launched last, read through the sgemm_cuda_last_* getters
	std::atomic<int> last_kernel{0};
	std::atomic<int> last_conv_fused{0};   // last convolution ran as an implicit GEMM (no column matrix)
	int conv_fusion = -1;            // -1 by rule, 0 never, 1 whenever the hard constraints allow
	std::atomic<int> last_repacked{0};     // last auto launch copied an operand to an aligned leading dimension first
	int sm_limit = 0;                // 0 = all SMs; otherwise K1's persistent grid is capped (leaves SMs to NCCL)
	std::atomic<unsigned long long> launches{0};
} g;

// serialises init / finish and the host-pointer entry points (they share the staging arena and the three streams); recursive
// because a host-pointer call initialises the library lazily while it already holds the lock
std::recursive_mutex g_mu;
std::mutex g_err_mu;    // the sticky error string
char g_err[512];
std::atomic<bool> g_has_err{false};

void set_error(const char *fmt, ...)
{
	std::lock_guard<std::mutex> lk(g_err_mu);
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	g_has_err = true;
	if (getenv("UGEMM_CUDA_VERBOSE")) fprintf(stderr, "ugemm_cuda: %s\n", g_err);
}

} // namespace
namespace ugemm { void report_error(const char *msg) { set_error("%s", msg); } }   // for the other translation units of the C ABI (shard.cu)
namespace {

#define CU_TRY(expr, what)                                                                   \
	do {                                                                                     \
		cudaError_t e__ = (expr);                                                            \
		if (e__ != cudaSuccess) {                                                            \
			set_error("%s failed: %s", what, cudaGetErrorString(e__));                       \
			return 1;                                                                        \
		}                                                                                    \
	} while (0)

// Every entry point starts here: lazy init, then bind the CALLING thread to the backend's device (a host thread other than
// the one that initialised the library starts on device 0, whatever device the backend lives on).
int ensure_init()
{
	if (!g.ready.load(std::memory_order_acquire) && sgemm_cuda_init(-1, 0)) return 1;
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess || cur != g.device) CU_TRY(cudaSetDevice(g.device), "cudaSetDevice");
	return 0;
}

size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

int upper(char c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// Validate BLAS arguments and map to the row-major Problem.  Column-major C = op(A) op(B) is the row-major
// product C^T = op(B)^T op(A)^T over the same buffers: swap (A,transA,M,lda) <-> (B,transB,N,ldb).
int normalise(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
              const float *B, int ldb, float beta, float *C, int ldc, Problem *p)
{
	major = (char)upper(major); ta = (char)upper(ta); tb = (char)upper(tb);
	if (major != 'R' && major != 'C') { set_error("major must be 'R' or 'C' (got 0x%02x)", major); return 1; }
	if ((ta != 'N' && ta != 'T') || (tb != 'N' && tb != 'T')) { set_error("transA/transB must be 'N' or 'T'"); return 1; }
	if (M < 0 || N < 0 || K < 0) { set_error("negative dimension M=%d N=%d K=%d", M, N, K); return 1; }
	if (major == 'C') {
		const float *tp = A; A = B; B = tp;
		int ti = lda; lda = ldb; ldb = ti;
		ti = M; M = N; N = ti;
		char tc = ta; ta = tb; tb = tc;
	}
	const int a_cols = ta == 'N' ? K : M, b_cols = tb == 'N' ? N : K;
	if (lda < (a_cols > 1 ? a_cols : 1) || ldb < (b_cols > 1 ? b_cols : 1) || ldc < (N > 1 ? N : 1)) {
		set_error("leading dimension too small (lda=%d ldb=%d ldc=%d for stored widths %d %d %d)", lda, ldb, ldc, a_cols, b_cols, N);
		return 1;
	}
	p->M = M; p->N = N; p->K = K; p->alpha = alpha; p->beta = beta;
	p->A = A; p->lda = lda; p->a_kmajor = (ta == 'N');
	p->B = B; p->ldb = ldb; p->b_kmajor = (tb == 'T');
	p->C = C; p->ldc = ldc;
	return 0;
}

// At least one full 128-wide tile side of tensor work; the other side may be as narrow as 48 (8 with K >= 512) once the problem is big
// enough to amortise K1's launch (its zero-filled tile columns cost nothing extra: a 200704 x 64 x 1152 product takes 0.50 ms on K1,
// 0.66 ms on K2; 200704 x 96: 0.50 vs 1.47 ms; at N = 32 K2's narrow tiles win, 0.40 vs 0.49 ms -- profiles/r1_skinny_k1_vs_k2.jsonl).
bool auto_prefers_k1(const Problem &p)
{
	if (!k1_eligible(p, nullptr) || p.K < 32) return false;
	if (p.M >= 128 && p.N >= 128) return true;
	if (!(p.M >= 128 || p.N >= 128) || (double)p.M * p.N * p.K < 67108864.0) return false;
	// round 2: the TS kernel neither multiplies nor promotes accumulator slices that hold no column of C, so from a narrow side of 8 it
	// beats K2's narrow tiles once K is long (>= 512): 200704 x 32 x 1152 0.24 vs 0.39 ms, 8192 x 16 x 4096 56 vs 232 us,
	// 4096 x 32 x 512 22 vs 40 us; with a short K the FFMA kernel keeps the very narrow shapes (100000 x 12 x 300: 74 vs 58 us)
	// -- tools/gpu_skinny3.py, profiles/r3i_skinny.json
	const int narrow = p.M < p.N ? p.M : p.N;
	return narrow >= 48 || (narrow >= 8 && p.K >= 512);
}

bool tma_ok(const float *ptr, long long ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0; }

// mode=auto, operand layout TMA cannot take (ld not a multiple of 4 or a misaligned base) but a problem big enough
// that the tensor cores pay for a copy: repack the offending operand(s) into a stream-ordered scratch buffer with an
// aligned leading dimension (one HBM pass over that operand, 2*4*rows*cols bytes, ~1 % of the GEMM time at the
// sizes this triggers for) and run K1 on the copy.  C needs no repacking: K1's epilogue stores through any ldc.
bool auto_wants_repack(const Problem &p)
{
	return p.batch <= 1 && !(tma_ok(p.A, p.lda) && tma_ok(p.B, p.ldb)) && p.M >= 256 && p.N >= 256 && p.K >= 64;
}

// returns 0 on success with *q the problem to launch and *scratch the buffer to cudaFreeAsync afterwards (or null);
// returns 1 if the scratch allocation is not available (caller falls back to K2, which is not an error)
int repack_for_tma(const Problem &p, cudaStream_t stream, Problem *q, void **scratch)
{
	*q = p; *scratch = nullptr;
	const long long a_lines = p.a_kmajor ? p.M : p.K, a_cols = p.a_kmajor ? p.K : p.M;
	const long long b_lines = p.b_kmajor ? p.N : p.K, b_cols = p.b_kmajor ? p.K : p.N;
	const bool ra = !tma_ok(p.A, p.lda), rb = !tma_ok(p.B, p.ldb);
	const long long lda2 = (a_cols + 3) / 4 * 4, ldb2 = (b_cols + 3) / 4 * 4;
	const size_t a_bytes = ra ? align_up_sz((size_t)a_lines * lda2 * 4, 256) : 0, b_bytes = rb ? (size_t)b_lines * ldb2 * 4 : 0;
	char *ws = nullptr;
	if (cudaMallocAsync(reinterpret_cast<void **>(&ws), a_bytes + b_bytes + 256, stream) != cudaSuccess) { cudaGetLastError(); return 1; }
	cudaError_t e = cudaSuccess;
	if (ra) {
		e = cudaMemcpy2DAsync(ws, (size_t)lda2 * 4, p.A, (size_t)p.lda * 4, (size_t)a_cols * 4, (size_t)a_lines, cudaMemcpyDeviceToDevice, stream);
		q->A = reinterpret_cast<const float *>(ws); q->lda = lda2;
	}
	if (rb && e == cudaSuccess) {
		e = cudaMemcpy2DAsync(ws + a_bytes, (size_t)ldb2 * 4, p.B, (size_t)p.ldb * 4, (size_t)b_cols * 4, (size_t)b_lines, cudaMemcpyDeviceToDevice, stream);
		q->B = reinterpret_cast<const float *>(ws + a_bytes); q->ldb = ldb2;
	}
	if (e != cudaSuccess) { cudaFreeAsync(ws, stream); cudaGetLastError(); return 1; }
	*scratch = ws;
	return 0;
}

// K1 launch with the backend's tuning.  While an SM limit is set (the sharded driver leaves SMs to NCCL, whose CTAs may delay
// some of K1's persistent pairs) the statically scheduled stream-K tail is switched off: the dynamic scheduler absorbs late
// pairs, a static schedule would wait for them.
cudaError_t launch_k1(const Problem &p, cudaStream_t stream)
{
	const bool limited = g.sm_limit > 0 && g.sm_limit < g.sm_count;
	K1Tuning t = g.tuning;
	// ... and no programmatic dependent launch either: the next product's CTAs would settle on the SMs left free for NCCL while this
	// product is still running, and the broadcasts would wait behind both (8 GPUs, NCCL transport: 42 -> 65 ms per step [measured])
	if (limited) t.flags |= 2048 | 262144;
	return launch_k1_3xtf32(p, t, stream, limited ? g.sm_limit : g.sm_count);
}

// device-pointer GEMM on `stream`
int run_dev(int mode, cudaStream_t stream, const Problem &p)
{
	// quick returns of the reference (sgemm_avx256.h:410)
	if (p.M == 0 || p.N == 0) return 0;
	if ((p.alpha == 0.f || p.K == 0) && p.beta == 1.f) return 0;
	if (p.alpha == 0.f || p.K == 0) {
		CU_TRY(launch_scale_c(p, stream), "scale_c launch");
		g.launches++;
		return 0;
	}
	int use = mode;
	if (mode == UGEMM_MODE_AUTO) use = auto_prefers_k1(p) ? UGEMM_MODE_3XTF32 : UGEMM_MODE_SIMT;
	if (mode == UGEMM_MODE_AUTO && use == UGEMM_MODE_SIMT && auto_wants_repack(p)) {
		Problem q; void *scratch = nullptr;
		if (repack_for_tma(p, stream, &q, &scratch) == 0 && k1_eligible(q, nullptr)) {
			cudaError_t e = launch_k1(q, stream);
			if (scratch) cudaFreeAsync(scratch, stream);
			CU_TRY(e, "K1 (3xTF32 tcgen05, repacked operands) launch");
			g.last_kernel = UGEMM_MODE_3XTF32; g.last_repacked = 1;
			g.launches++;
			return 0;
		}
		if (scratch) cudaFreeAsync(scratch, stream);
	}
	g.last_repacked = 0;
	if (use == UGEMM_MODE_3XTF32) {
		const char *why = nullptr;
		if (!k1_eligible(p, &why)) { set_error("3xTF32 kernel not applicable: %s", why); return 1; }
		CU_TRY(launch_k1(p, stream), "K1 (3xTF32 tcgen05) launch");
	} else if (use == UGEMM_MODE_SIMT) {
		CU_TRY(launch_k2_simt(p, stream, g.sm_count), "K2 (SIMT FFMA) launch");
	} else {
		set_error("unknown mode %d", mode);
		return 1;
	}
	g.last_kernel = use;
	g.launches++;
	return 0;
}

int ensure_arena(size_t bytes)
{
	if (bytes <= g.arena_bytes) return 0;
	if (g.arena) cudaFree(g.arena);
	g.arena = nullptr; g.arena_bytes = 0;
	size_t want = bytes + (bytes >> 3) + (1u << 20);
	CU_TRY(cudaMalloc(&g.arena, want), "arena cudaMalloc");
	g.arena_bytes = want;
	return 0;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

cudaError_t copy2d(float *dst, const float *src, long long ld, long long lines, long long cols, cudaMemcpyKind kind, cudaStream_t st)
{
	if (lines <= 0 || cols <= 0) return cudaSuccess;
	if (ld == cols) return cudaMemcpyAsync(dst, src, (size_t)lines * cols * 4, kind, st);
	return cudaMemcpy2DAsync(dst, (size_t)ld * 4, src, (size_t)ld * 4, (size_t)cols * 4, (size_t)lines, kind, st);
}

// Large problems: the reference's OpenCL path uploads everything, runs, downloads (sgemm_ocl2.h:175-217) and its
// timed region is dominated by the two PCIe transfers.  Here op(B) goes up first, then row panels of op(A) (and of C
// when beta != 0) stream up on one stream while the GEMM of the previous panel runs on a second and the finished C
// panel streams down on a third (PCIe is full duplex).  Same device layout and leading dimensions as the one-shot
// path.  The kernel is chosen ONCE, on the whole problem, and forced for every panel (a ragged last panel must not fall to
// another kernel with another rounding than the rest of C), and a last panel of fewer than 256 rows is merged into the
// one before it.  Host buffers should be pinned (ugemm_cuda_malloc_host) for the
// copies to be truly asynchronous; pageable memory still works, just without the overlap.
// On a failure in the middle of the pipeline the panels already downloaded stay in the caller's C (partial result; the
// error is reported): a CUDA failure at that point leaves the context unusable anyway.
void run_host_pipelined(int mode, const Problem &p, float *dA, float *dB, float *dC)
{
	const long long b_lines = p.b_kmajor ? p.N : p.K, b_cols = p.b_kmajor ? p.K : p.N;
	int panels = (int)((p.M + 511) / 512);     // small panels shorten the un-overlapped tail (last GEMM + last download)
	if (const char *e = getenv("UGEMM_CUDA_PANELS")) panels = atoi(e) > 0 ? atoi(e) : panels;
	if (panels > 16) panels = 16;
	long long mb = ((p.M + panels - 1) / panels + 255) / 256 * 256;
	panels = (int)((p.M + mb - 1) / mb);
	if (panels > 1 && p.M - (long long)(panels - 1) * mb < 256) panels--;      // the ragged tail joins the previous panel
	if (mode == UGEMM_MODE_AUTO) {
		Problem whole = p; whole.A = dA; whole.B = dB; whole.C = dC;
		// rule (1) -> K1 for every panel; rule (2) (repack) stays per panel in auto mode: every panel then has >= 256 rows,
		// so all of them repack and run K1; rule (3) -> K2 for every panel
		if (auto_prefers_k1(whole)) mode = UGEMM_MODE_3XTF32;
		else if (!auto_wants_repack(whole)) mode = UGEMM_MODE_SIMT;
	}
	cudaStream_t s_up = g.stream_h2d, s_cmp = g.stream, s_dn = g.stream_d2h;
	cudaError_t e = copy2d(dB, p.B, p.ldb, b_lines, b_cols, cudaMemcpyHostToDevice, s_up);
	int rc = 0;
	for (int i = 0; i < panels && e == cudaSuccess && !rc; i++) {
		const long long m0 = i * mb, mm = (i == panels - 1) ? p.M - m0 : mb;
		// panel i of op(A): rows m0.. of an M x K array (k-major) or columns m0.. of a K x M array
		if (p.a_kmajor) e = copy2d(dA + m0 * p.lda, p.A + m0 * p.lda, p.lda, mm, p.K, cudaMemcpyHostToDevice, s_up);
		else e = cudaMemcpy2DAsync(dA + m0, (size_t)p.lda * 4, p.A + m0, (size_t)p.lda * 4, (size_t)mm * 4, (size_t)p.K, cudaMemcpyHostToDevice, s_up);
		if (e == cudaSuccess && p.beta != 0.f)
			e = copy2d(dC + m0 * p.ldc, p.C + m0 * p.ldc, p.ldc, mm, p.N, cudaMemcpyHostToDevice, s_up);
		if (e != cudaSuccess) break;
		cudaEventRecord(g.ev_up[i], s_up);
		cudaStreamWaitEvent(s_cmp, g.ev_up[i], 0);
		Problem d = p;
		d.M = (int)mm;
		d.A = p.a_kmajor ? dA + m0 * p.lda : dA + m0;
		d.B = dB;
		d.C = dC + m0 * p.ldc;
		rc = run_dev(mode, s_cmp, d);
		if (rc) break;
		cudaEventRecord(g.ev_done[i], s_cmp);
		cudaStreamWaitEvent(s_dn, g.ev_done[i], 0);
		e = copy2d(p.C + m0 * p.ldc, dC + m0 * p.ldc, p.ldc, mm, p.N, cudaMemcpyDeviceToHost, s_dn);
	}
	cudaError_t e1 = cudaStreamSynchronize(s_up), e2 = cudaStreamSynchronize(s_cmp), e3 = cudaStreamSynchronize(s_dn);
	if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
	if (e != cudaSuccess && !rc) {
		const unsigned *dg = k1_diag_host();
		if (dg && dg[0]) set_error("pipelined GEMM failed: %s (K1 watchdog code %u, block %u, thread %u)", cudaGetErrorString(e), dg[0], dg[1], dg[2]);
		else set_error("pipelined GEMM failed: %s", cudaGetErrorString(e));
	}
}

// Host-pointer GEMM: stage operands into the arena, run, copy C back.  Blocking, like sgemm_ocl.
void run_host(int mode, char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
              const float *B, int ldb, float beta, float *C, int ldc)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &p)) return;
	if (p.M == 0 || p.N == 0) return;
	if ((p.alpha == 0.f || p.K == 0) && p.beta == 1.f) return;

	const bool need_ab = !(p.alpha == 0.f || p.K == 0);
	const long long a_lines = p.a_kmajor ? p.M : p.K, a_cols = p.a_kmajor ? p.K : p.M;
	const long long b_lines = p.b_kmajor ? p.N : p.K, b_cols = p.b_kmajor ? p.K : p.N;
	// device copies keep the caller's leading dimensions (so eligibility and alignment are the caller's)
	const size_t a_bytes = need_ab ? (size_t)((a_lines - 1) * p.lda + a_cols) * 4 : 0;
	const size_t b_bytes = need_ab ? (size_t)((b_lines - 1) * p.ldb + b_cols) * 4 : 0;
	const size_t c_bytes = (size_t)((long long)(p.M - 1) * p.ldc + p.N) * 4;
	const size_t offA = 0, offB = align_up(offA + a_bytes, 256), offC = align_up(offB + b_bytes, 256);
	if (ensure_arena(offC + c_bytes)) return;
	float *dA = reinterpret_cast<float *>(g.arena + offA), *dB = reinterpret_cast<float *>(g.arena + offB);
	float *dC = reinterpret_cast<float *>(g.arena + offC);

	// the panel pipeline only pays when the transfers are long (>= 64 MB in flight); small problems go up, run and come back whole
	if (need_ab && p.M >= 2048 && a_bytes + b_bytes + c_bytes >= (64u << 20) && !getenv("UGEMM_CUDA_NO_PIPELINE")) {
		if (mode == UGEMM_MODE_3XTF32) {
			Problem chk = p; chk.A = dA; chk.B = dB; chk.C = dC;
			const char *why = nullptr;
			if (!k1_eligible(chk, &why)) { set_error("3xTF32 kernel not applicable: %s", why); return; }
		}
		run_host_pipelined(mode, p, dA, dB, dC);
		return;
	}

	cudaError_t e = cudaSuccess;
	auto up2d = [&](float *d, const float *h, long long ld, long long lines, long long cols) {
		if (e != cudaSuccess || lines <= 0 || cols <= 0) return;
		if (ld == cols) e = cudaMemcpyAsync(d, h, (size_t)lines * cols * 4, cudaMemcpyHostToDevice, g.stream);
		else e = cudaMemcpy2DAsync(d, (size_t)ld * 4, h, (size_t)ld * 4, (size_t)cols * 4, (size_t)lines, cudaMemcpyHostToDevice, g.stream);
	};
	if (need_ab) {
		up2d(dA, p.A, p.lda, a_lines, a_cols);
		up2d(dB, p.B, p.ldb, b_lines, b_cols);
	}
	if (p.beta != 0.f) up2d(dC, p.C, p.ldc, p.M, p.N);   // C is uploaded only when it is read (sgemm_ocl2.h:177)
	if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); return; }

	Problem d = p;
	d.A = dA; d.B = dB; d.C = dC;
	if (mode == UGEMM_MODE_3XTF32) {
		// eligibility is judged on the caller's layout; the arena copy preserves ld and 256-B alignment
		const char *why = nullptr;
		if (!k1_eligible(d, &why)) { set_error("3xTF32 kernel not applicable: %s", why); return; }
	}
	if (run_dev(mode, g.stream, d)) return;
	// only the M x N region comes back: ld padding on the host is never written
	if (p.ldc == p.N) e = cudaMemcpyAsync(p.C, dC, (size_t)p.M * p.N * 4, cudaMemcpyDeviceToHost, g.stream);
	else e = cudaMemcpy2DAsync(p.C, (size_t)p.ldc * 4, dC, (size_t)p.ldc * 4, (size_t)p.N * 4, (size_t)p.M, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) {
		const unsigned *dg = k1_diag_host();
		if (dg && dg[0]) set_error("GEMM failed: %s (K1 watchdog code %u, block %u, thread %u)", cudaGetErrorString(e), dg[0], dg[1], dg[2]);
		else set_error("GEMM failed: %s", cudaGetErrorString(e));
	}
}

__global__ void fill_uniform_kernel(float *x, size_t n, unsigned long long base, float lo, float span)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) x[i] = uniform_at(base, i, lo, span);
}

// dst[r*ld + c] = stream element (offset + r*gld + c): a window of a larger row-major matrix
__global__ void fill_uniform_2d_kernel(float *x, long long rows, long long cols, long long ld, unsigned long long base,
                                       unsigned long long offset, unsigned long long gld, float lo, float span)
{
	const long long total = rows * cols;
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (; i < total; i += stride) {
		const long long r = i / cols, c = i - r * cols;
		x[r * ld + c] = uniform_at(base, offset + (unsigned long long)r * gld + (unsigned long long)c, lo, span);
	}
}

} // namespace

extern "C" {

int sgemm_cuda_init(int device, size_t arena_bytes)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (device < 0) {
		const char *env = getenv("UGEMM_CUDA_DEVICE");
		device = env ? atoi(env) : (g.ready ? g.device : 0);
	}
	if (g.ready) {
		// one backend, one device (like the reference's process-global cl context, ocl.h:141-193): a second init on ANOTHER
		// device is an error, not a silent no-op that leaves the caller's buffers on the wrong GPU
		if (device != g.device) { set_error("sgemm_cuda_init(%d): the backend is already initialised on device %d; call sgemm_cuda_finish first", device, g.device); return 1; }
		if (arena_bytes) return ensure_arena(arena_bytes);
		return 0;
	}
	int count = 0;
	CU_TRY(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
	if (device >= count) { set_error("device %d requested but only %d visible", device, count); return 1; }
	CU_TRY(cudaSetDevice(device), "cudaSetDevice");
	cudaDeviceProp prop;
	CU_TRY(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
	if (prop.major != 10) {
		set_error("device %d (%s) is sm_%d%d; this backend is built for sm_100a only", device, prop.name, prop.major, prop.minor);
		return 1;
	}
	g.device = device;
	g.sm_count = prop.multiProcessorCount;
	g.hbm_bytes = prop.totalGlobalMem;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
	g.clock_khz = khz;
	strncpy(g.name, prop.name, sizeof g.name - 1);
	CU_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking), "cudaStreamCreate");
	CU_TRY(cudaStreamCreateWithFlags(&g.stream_d2h, cudaStreamNonBlocking), "cudaStreamCreate");
	CU_TRY(cudaStreamCreateWithFlags(&g.stream_h2d, cudaStreamNonBlocking), "cudaStreamCreate");
	for (int i = 0; i < 16; i++) {
		CU_TRY(cudaEventCreateWithFlags(&g.ev_up[i], cudaEventDisableTiming), "cudaEventCreate");
		CU_TRY(cudaEventCreateWithFlags(&g.ev_done[i], cudaEventDisableTiming), "cudaEventCreate");
	}
	{   // keep stream-ordered scratch (the repack path) cached in the pool instead of returning it to the OS at every sync
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			unsigned long long keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
	}
	if (const char *f = getenv("UGEMM_K1_FLAGS")) {
		// overrides the default (bit0 = A collector on).  Bits 1-6 are ablation / profiling switches that make K1's RESULTS WRONG
		// (common.cuh): a stray environment variable must not silently corrupt a production run, so they need an explicit opt-in.
		const int flags = atoi(f), unsafe = flags & (2 | 4 | 8 | 16 | 32 | 64 | 65536);
		const char *opt = getenv("UGEMM_K1_ABLATION");
		if (unsafe && !(opt && atoi(opt) == 1)) {
			set_error("UGEMM_K1_FLAGS=%d sets ablation bits 0x%x that corrupt results; set UGEMM_K1_ABLATION=1 to allow them (bottleneck analysis only)", flags, unsafe);
			return 1;
		}
		g.tuning.flags = flags;
	}
	g.ready.store(true, std::memory_order_release);
	if (arena_bytes && ensure_arena(arena_bytes)) { g.ready = false; return 1; }
	return 0;
}

void sgemm_cuda_finish(void)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (!g.ready) return;
	cudaSetDevice(g.device);
	cudaStreamSynchronize(g.stream);
	if (g.arena) cudaFree(g.arena);
	g.arena = nullptr; g.arena_bytes = 0;
	cudaStreamDestroy(g.stream);
	cudaStreamDestroy(g.stream_d2h);
	cudaStreamDestroy(g.stream_h2d);
	for (int i = 0; i < 16; i++) { cudaEventDestroy(g.ev_up[i]); cudaEventDestroy(g.ev_done[i]); }
	g.stream = g.stream_d2h = g.stream_h2d = nullptr;
	g.ready = false;
}

void sgemm_cuda(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
                const float *B, int ldb, float beta, float *C, int ldc)
{ run_host(UGEMM_MODE_AUTO, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); }

void sgemm_cuda_3xtf32(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
                       const float *B, int ldb, float beta, float *C, int ldc)
{ run_host(UGEMM_MODE_3XTF32, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); }

void sgemm_cuda_simt(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
                     const float *B, int ldb, float beta, float *C, int ldc)
{ run_host(UGEMM_MODE_SIMT, major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); }

int sgemm_cuda_dev(int mode, void *stream, char major, char ta, char tb, int M, int N, int K, float alpha,
                   const float *dA, int lda, const float *dB, int ldb, float beta, float *dC, int ldc)
{
	if (ensure_init()) return 1;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &p)) return 1;
	return run_dev(mode, stream ? static_cast<cudaStream_t>(stream) : g.stream, p);
}

// Strided batch: `batch` independent problems of identical shape, instance b at A + b*strideA, B + b*strideB,
// C + b*strideC -- the stacked layout test_sgemm walks one call at a time (check_sgemm.c:111-124), done in ONE launch
// so that small instances still fill the machine.  Device pointers, asynchronous.
int sgemm_cuda_batched_dev(int mode, void *stream, char major, char ta, char tb, int M, int N, int K, float alpha,
                           const float *dA, int lda, long long strideA, const float *dB, int ldb, long long strideB,
                           float beta, float *dC, int ldc, long long strideC, int batch)
{
	if (ensure_init()) return 1;
	if (batch < 0) { set_error("negative batch count"); return 1; }
	if (batch == 0) return 0;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &p)) return 1;
	const bool swapped = (major == 'C' || major == 'c');
	p.batch = batch;
	p.strideA = swapped ? strideB : strideA;     // normalise() swapped the operands for column-major
	p.strideB = swapped ? strideA : strideB;
	p.strideC = strideC;
	if (batch > 1 && (p.strideA < 0 || p.strideB < 0 || p.strideC < (long long)(p.M - 1) * p.ldc + p.N)) {
		set_error("batch strides must be non-negative and strideC must not make instances of C overlap");
		return 1;
	}
	return run_dev(mode, stream ? static_cast<cudaStream_t>(stream) : g.stream, p);
}

// host-pointer version: whole stacked buffers go up, one launch, the stacked C comes back (blocking)
void sgemm_cuda_batched(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda, long long strideA,
                        const float *B, int ldb, long long strideB, float beta, float *C, int ldc, long long strideC, int batch)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	if (batch < 0) { set_error("negative batch count"); return; }
	if (batch == 0) return;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &p)) return;
	if (p.M == 0 || p.N == 0) return;
	if ((p.alpha == 0.f || p.K == 0) && p.beta == 1.f) return;          // quick return of the reference (sgemm_avx256.h:410)
	const bool swapped = (major == 'C' || major == 'c');
	const long long sA = swapped ? strideB : strideA, sB = swapped ? strideA : strideB;
	// same stride rules as the device-pointer entry point, checked BEFORE any size arithmetic uses them
	if (batch > 1 && (sA < 0 || sB < 0 || strideC < (long long)(p.M - 1) * p.ldc + p.N)) {
		set_error("batch strides must be non-negative and strideC must not make instances of C overlap");
		return;
	}
	// alpha == 0 or K == 0: A and B are never read (the caller may pass NULL), only C <- beta*C happens -- like run_host
	const bool need_ab = !(p.alpha == 0.f || p.K == 0);
	const long long a_lines = p.a_kmajor ? p.M : p.K, a_cols = p.a_kmajor ? p.K : p.M;
	const long long b_lines = p.b_kmajor ? p.N : p.K, b_cols = p.b_kmajor ? p.K : p.N;
	const size_t a_n = need_ab ? (size_t)((batch - 1) * sA + (a_lines - 1) * p.lda + a_cols) : 0;
	const size_t b_n = need_ab ? (size_t)((batch - 1) * sB + (b_lines - 1) * p.ldb + b_cols) : 0;
	const size_t c_n = (size_t)((batch - 1) * strideC + (long long)(p.M - 1) * p.ldc + p.N);
	const size_t offB = align_up(a_n * 4, 256), offC = align_up(offB + b_n * 4, 256);
	if (ensure_arena(offC + c_n * 4)) return;
	float *dA = reinterpret_cast<float *>(g.arena), *dB = reinterpret_cast<float *>(g.arena + offB), *dC = reinterpret_cast<float *>(g.arena + offC);
	cudaError_t e = cudaSuccess;
	if (need_ab) {
		e = cudaMemcpyAsync(dA, p.A, a_n * 4, cudaMemcpyHostToDevice, g.stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dB, p.B, b_n * 4, cudaMemcpyHostToDevice, g.stream);
	}
	// C travels whole (padding included, so the download restores it bit for bit); needed up front when beta != 0
	if (e == cudaSuccess) e = cudaMemcpyAsync(dC, p.C, c_n * 4, cudaMemcpyHostToDevice, g.stream);
	if (e != cudaSuccess) { set_error("batched H2D failed: %s", cudaGetErrorString(e)); return; }
	Problem d = p;
	d.A = dA; d.B = dB; d.C = dC; d.batch = batch; d.strideA = sA; d.strideB = sB; d.strideC = strideC;
	if (run_dev(UGEMM_MODE_AUTO, g.stream, d)) return;
	e = cudaMemcpyAsync(p.C, dC, c_n * 4, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("batched GEMM failed: %s", cudaGetErrorString(e));
}

int sgemm_cuda_k1_eligible(char major, char ta, char tb, int M, int N, int K, const float *dA, int lda,
                           const float *dB, int ldb, const float *dC, int ldc)
{
	if (ensure_init()) return 0;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, 1.f, dA, lda, dB, ldb, 0.f, const_cast<float *>(dC), ldc, &p)) {
		sgemm_cuda_clear_error();
		return 0;
	}
	return auto_prefers_k1(p) ? 1 : 0;
}

int sgemm_cuda_time_dev(int mode, int iters, int warmup, char major, char ta, char tb, int M, int N, int K,
                        float alpha, const float *dA, int lda, const float *dB, int ldb, float beta, float *dC,
                        int ldc, float *ms_avg, float *ms_min, float *ms_total)
{
	if (ensure_init()) return 1;
	if (iters < 1) iters = 1;
	if (iters > 4096) iters = 4096;
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &p)) return 1;
	for (int i = 0; i < warmup; i++)
		if (run_dev(mode, g.stream, p)) return 1;
	CU_TRY(cudaStreamSynchronize(g.stream), "warm-up sync");
	// one event pair per launch, all enqueued back to back (no host sync in between): per-launch durations
	// AND the whole-region time come from the same run
	cudaEvent_t *ev = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * 2 * (size_t)iters);
	if (!ev) { set_error("out of host memory"); return 1; }
	int made = 0, rc = 0;
	for (; made < 2 * iters; made++)
		if (cudaEventCreate(&ev[made]) != cudaSuccess) { set_error("cudaEventCreate failed"); rc = 1; break; }
	for (int i = 0; i < iters && !rc; i++) {
		cudaEventRecord(ev[2 * i], g.stream);
		rc = run_dev(mode, g.stream, p);
		cudaEventRecord(ev[2 * i + 1], g.stream);
	}
	if (!rc) {
		cudaError_t e = cudaStreamSynchronize(g.stream);
		if (e != cudaSuccess) {
			const unsigned *dg = k1_diag_host();
			if (dg && dg[0]) set_error("timed launch failed: %s (K1 watchdog code %u, block %u, thread %u)", cudaGetErrorString(e), dg[0], dg[1], dg[2]);
			else set_error("timed launch failed: %s", cudaGetErrorString(e));
			rc = 1;
		}
	}
	if (!rc) {
		float total = 0.f, best = 1e30f, span = 0.f;
		for (int i = 0; i < iters; i++) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
			total += ms;
			if (ms < best) best = ms;
		}
		cudaEventElapsedTime(&span, ev[0], ev[2 * iters - 1]);
		if (ms_avg) *ms_avg = total / iters;
		if (ms_min) *ms_min = best;
		if (ms_total) *ms_total = span;
	}
	for (int i = 0; i < made; i++) cudaEventDestroy(ev[i]);
	free(ev);
	return rc;
}

const char *sgemm_cuda_last_error(void) { return g_has_err ? g_err : nullptr; }
void sgemm_cuda_clear_error(void) { std::lock_guard<std::mutex> lk(g_err_mu); g_has_err = false; g_err[0] = 0; }
int sgemm_cuda_last_kernel(void) { return g.last_kernel; }
int sgemm_cuda_last_repacked(void) { return g.last_repacked; }
int sgemm_cuda_last_conv_fused(void) { return g.last_conv_fused; }
void sgemm_cuda_set_conv_fusion(int mode) { g.conv_fusion = mode < 0 ? -1 : (mode > 0 ? 1 : 0); }
unsigned long long sgemm_cuda_launch_count(void) { return g.launches; }

int ugemm_cuda_device_info(int *sm_count, int *sm_clock_khz, size_t *hbm_bytes, char *name, int name_len)
{
	if (ensure_init()) return 1;
	if (sm_count) *sm_count = g.sm_count;
	if (sm_clock_khz) *sm_clock_khz = g.clock_khz;
	if (hbm_bytes) *hbm_bytes = g.hbm_bytes;
	if (name && name_len > 0) { strncpy(name, g.name, (size_t)name_len - 1); name[name_len - 1] = 0; }
	return 0;
}

void sgemm_cuda_set_k1_tuning(int kc_blocks, int split, int cta_group)
{
	if (kc_blocks >= 0) g.tuning.kc_blocks = kc_blocks;
	if (split >= 0) g.tuning.split = split ? 1 : 0;
	if (cta_group >= 0 && cta_group <= 2) g.tuning.cta_group = cta_group;   // 0 = choose by problem size
}

// The schedule K1 would use for a dense M x N x K problem (`batch` instances) on `sm_count` SMs (0: 148, a B200) with the current
// tuning: pure host arithmetic, no GPU needed.  out[0..11] = cta_group, tile_m, tile_n, tiles_m, tiles_n, k-blocks per tile,
// promotion interval (k-blocks), whole tiles, stream-K tail tiles, promotion chunks per tile, chunks per tail range, work items.
int sgemm_cuda_k1_plan(int M, int N, int K, int batch, int sm_count, int *out12)
{
	if (M < 1 || N < 1 || K < 1 || !out12) { set_error("sgemm_cuda_k1_plan: bad arguments"); return 1; }
	k1_plan(M, N, K, batch > 0 ? batch : 1, g.tuning, sm_count > 0 ? sm_count : 148, out12);
	return 0;
}
// Segment h (0 or 1) of work item `item` of such a schedule: out[0..3] = tile, first k-block, end k-block, workspace slot (-1 for a
// whole tile); an absent segment has end <= first.  The same function the kernel's roles decode their work with.
int sgemm_cuda_k1_plan_item(const int *plan12, int item, int h, int *out4)
{
	if (!plan12 || !out4 || item < 0 || item >= plan12[11] || h < 0 || h > 1) { set_error("sgemm_cuda_k1_plan_item: bad arguments"); return 1; }
	k1_plan_item(item, h, plan12[7], plan12[8], plan12[9], plan12[10], plan12[6], plan12[5], out4);
	return 0;
}

void sgemm_cuda_set_k1_variant(int variant)
{
	if (variant == 1) g.tuning.flags |= 32768;        // round-1 SS kernel (both operands from shared memory)
	else if (variant == 0) g.tuning.flags &= ~32768;  // TS kernel (A in tensor memory): the default
}

void sgemm_cuda_set_sm_limit(int sms) { g.sm_limit = sms > 0 ? sms : 0; }

void *ugemm_cuda_malloc(size_t bytes)
{
	if (ensure_init()) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
	return p;
}
void ugemm_cuda_free(void *d) { if (d) cudaFree(d); }
void *ugemm_cuda_malloc_host(size_t bytes)
{
	if (ensure_init()) return nullptr;
	void *p = nullptr;
	cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
	return p;
}
void ugemm_cuda_free_host(void *h) { if (h) cudaFreeHost(h); }
int ugemm_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
	if (ensure_init()) return 1;
	CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g.stream), "H2D copy");
	CU_TRY(cudaStreamSynchronize(g.stream), "H2D sync");
	return 0;
}
int ugemm_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
	if (ensure_init()) return 1;
	CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.stream), "D2H copy");
	CU_TRY(cudaStreamSynchronize(g.stream), "D2H sync");
	return 0;
}
int ugemm_cuda_memcpy_async(void *dst, const void *src, size_t bytes, void *stream)
{
	if (ensure_init()) return 1;
	CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream ? static_cast<cudaStream_t>(stream) : g.stream), "async copy");
	return 0;
}
int ugemm_cuda_ipc_export(void *dptr, void *handle64)
{
	if (ensure_init()) return 1;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	CU_TRY(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), dptr), "cudaIpcGetMemHandle");
	return 0;
}
void *ugemm_cuda_ipc_import(const void *handle64)
{
	if (ensure_init()) return nullptr;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof h);
	void *p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); return nullptr; }
	return p;
}
int ugemm_cuda_ipc_close(void *p)
{
	if (!p) return 0;
	CU_TRY(cudaIpcCloseMemHandle(p), "cudaIpcCloseMemHandle");
	return 0;
}

int ugemm_cuda_sync(void)
{
	if (ensure_init()) return 1;
	cudaError_t e = cudaStreamSynchronize(g.stream);
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		const unsigned *dg = k1_diag_host();
		if (dg && dg[0]) set_error("sync failed: %s (K1 watchdog code %u, block %u, thread %u)", cudaGetErrorString(e), dg[0], dg[1], dg[2]);
		else set_error("sync failed: %s", cudaGetErrorString(e));
		return 1;
	}
	return 0;
}

void ugemm_fill_uniform_host(float *x, size_t n, uint64_t seed, float lo, float hi)
{
	const unsigned long long base = seed * 0x9E3779B97F4A7C15ull;
	const float span = hi - lo;
	for (size_t i = 0; i < n; i++) x[i] = uniform_at(base, i, lo, span);
}

int ugemm_fill_uniform_dev(float *dx, size_t n, uint64_t seed, float lo, float hi, void *stream)
{
	if (ensure_init()) return 1;
	if (n == 0) return 0;
	cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : g.stream;
	size_t blocks = (n + 255) / 256;
	const size_t cap = (size_t)g.sm_count * 32;
	if (blocks > cap) blocks = cap;
	fill_uniform_kernel<<<(unsigned)blocks, 256, 0, s>>>(dx, n, seed * 0x9E3779B97F4A7C15ull, lo, hi - lo);
	CU_TRY(cudaGetLastError(), "fill_uniform launch");
	g.launches++;
	return 0;
}

void ugemm_fill_uniform_host_2d(float *x, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                                uint64_t gld, float lo, float hi)
{
	const unsigned long long base = seed * 0x9E3779B97F4A7C15ull;
	const float span = hi - lo;
	for (size_t r = 0; r < rows; r++)
		for (size_t c = 0; c < cols; c++) x[r * ld + c] = uniform_at(base, offset + r * gld + c, lo, span);
}

int ugemm_fill_uniform_dev_2d(float *dx, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                              uint64_t gld, float lo, float hi, void *stream)
{
	if (ensure_init()) return 1;
	if (rows == 0 || cols == 0) return 0;
	cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : g.stream;
	size_t blocks = (rows * cols + 255) / 256;
	const size_t cap = (size_t)g.sm_count * 32;
	if (blocks > cap) blocks = cap;
	fill_uniform_2d_kernel<<<(unsigned)blocks, 256, 0, s>>>(dx, (long long)rows, (long long)cols, (long long)ld,
	                                                        seed * 0x9E3779B97F4A7C15ull, offset, gld, lo, hi - lo);
	CU_TRY(cudaGetLastError(), "fill_uniform_2d launch");
	g.launches++;
	return 0;
}

// ---- convolution callers of the GEMM -------------------------------------------------
int im2col_cuda_dev(const float *d_im, int channels, int height, int width, int k, int pad, int stride, float *d_col, void *stream)
{
	if (ensure_init()) return 1;
	if (channels < 0 || height <= 0 || width <= 0 || k <= 0 || stride <= 0 || pad < 0 || height + 2 * pad < k || width + 2 * pad < k) {
		set_error("im2col: bad geometry (C=%d H=%d W=%d k=%d pad=%d stride=%d)", channels, height, width, k, pad, stride);
		return 1;
	}
	CU_TRY(launch_im2col(d_im, channels, height, width, k, pad, stride, d_col, stream ? static_cast<cudaStream_t>(stream) : g.stream), "im2col launch");
	g.launches++;
	return 0;
}

// Implicit-GEMM convolution on K1 (strides 1..8).  Under the automatic rule an efficiency bound applies: the GEMM is computed on
// a column index padded to a multiple of 32 per output row and on a channel count padded to a multiple of 32, and the padded
// work must stay within 30 % of the real work -- otherwise materialising the column matrix and running the dense GEMM is cheaper.
static bool conv_fusable(int mode, int nimg, int ich, int w, int h, int k, int pad, int stride, int ch)
{
	if (mode == UGEMM_MODE_SIMT || g.conv_fusion == 0 || stride < 1 || stride > 8) return false;   // TMA box <= 256 elements: 32*stride
	const long long ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
	if (ho < 1 || wo < 1 || ho * ((wo + 31) / 32 * 32) > 0x7fffffffLL) return false;
	if (g.conv_fusion > 0) return true;
	const long long wp = (wo + 31) / 32 * 32, ichp = (ich + 31) / 32 * 32;
	return ch >= 64 && ho * wo * nimg >= 256 && (double)(wp * ichp) <= 1.3 * (double)(wo * ich);
}

static int conv_fused_dev(cudaStream_t stream, const float *d_inputs, int nimg, int ich, int w, int h, const float *d_weights, int k, int pad,
                          int stride, float *d_outputs, int ch, const float *d_bias, float slope)
{
	ConvProblem c;
	c.nimg = nimg; c.ich = ich; c.h = h; c.w = w; c.ichp = (ich + 31) / 32 * 32; c.cs = (ich + 3) / 4 * 4;
	c.k = k; c.pad = pad; c.stride = stride; c.ho = (h + 2 * pad - k) / stride + 1; c.wo = (w + 2 * pad - k) / stride + 1; c.ch = ch;
	c.out = d_outputs; c.bias = d_bias; c.slope = slope;
	// scratch: channels-last copy of the images (image-sized, k*k times smaller than the column matrix) + repacked weights
	const size_t hwc_bytes = align_up_sz((size_t)nimg * h * w * c.cs * sizeof(float), 256), w_bytes = (size_t)ch * k * k * c.ichp * sizeof(float);
	char *ws = nullptr;
	CU_TRY(cudaMallocAsync(reinterpret_cast<void **>(&ws), hwc_bytes + w_bytes, stream), "convolution scratch");
	float *hwc = reinterpret_cast<float *>(ws), *wr = reinterpret_cast<float *>(ws + hwc_bytes);
	c.in_hwc = hwc; c.wgt_kkc = wr;
	cudaError_t e = launch_chw_to_hwc(d_inputs, nimg, ich, h, w, c.cs, hwc, stream);
	if (e == cudaSuccess) e = launch_conv_weight_repack(d_weights, ch, ich, k, c.ichp, wr, stream);
	if (e == cudaSuccess) e = launch_k1_conv(c, g.tuning, stream, (g.sm_limit > 0 && g.sm_limit < g.sm_count) ? g.sm_limit : g.sm_count);
	cudaFreeAsync(ws, stream);
	CU_TRY(e, "K1 implicit-GEMM convolution launch");
	g.last_kernel = UGEMM_MODE_3XTF32; g.last_conv_fused = 1; g.last_repacked = 0;
	g.launches += 3;
	return 0;
}

static int conv_check(int nimg, int ich, int w, int h, int k, int pad, int stride, int ch)
{
	if (nimg < 0 || ich <= 0 || ch <= 0 || k <= 0 || stride <= 0 || pad < 0 || h <= 0 || w <= 0 || h + 2 * pad < k || w + 2 * pad < k) {
		set_error("convolution: bad geometry (images=%d C=%d H=%d W=%d k=%d pad=%d stride=%d ch=%d)", nimg, ich, h, w, k, pad, stride, ch);
		return 1;
	}
	return 0;
}

int convolution_cuda_batched_dev(int mode, void *stream, const float *d_inputs, int nimg, int ich, int w, int h, const float *d_weights, int k,
                                 int pad, int stride, float *d_outputs, int ch, const float *d_bias, float slope, float *d_workspace)
{
	if (ensure_init()) return 1;
	if (conv_check(nimg, ich, w, h, k, pad, stride, ch)) return 1;
	if (nimg == 0) return 0;
	cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
	if (conv_fusable(mode, nimg, ich, w, h, k, pad, stride, ch))
		return conv_fused_dev(st, d_inputs, nimg, ich, w, h, d_weights, k, pad, stride, d_outputs, ch, d_bias, slope);
	if (!d_workspace) { set_error("convolution: this geometry takes the im2col path and needs d_workspace (ich*k*k*Ho*Wo floats)"); return 1; }
	const long long ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
	for (int i = 0; i < nimg; i++)      // the column matrix of one image at a time through the same workspace (stream order keeps it safe)
		if (convolution_cuda_dev(mode, st, d_inputs + (size_t)i * ich * h * w, ich, w, h, d_weights, k, pad, stride,
		                         d_outputs + (size_t)i * ch * ho * wo, ch, d_bias, slope, d_workspace)) return 1;
	return 0;
}

int convolution_cuda_dev(int mode, void *stream, const float *d_inputs, int ich, int w, int h, const float *d_weights, int k,
                         int pad, int stride, float *d_outputs, int ch, const float *d_bias, float slope, float *d_workspace)
{
	if (ensure_init()) return 1;
	if (conv_check(1, ich, w, h, k, pad, stride, ch)) return 1;
	g.last_conv_fused = 0;
	if (conv_fusable(mode, 1, ich, w, h, k, pad, stride, ch))
		return conv_fused_dev(stream ? static_cast<cudaStream_t>(stream) : g.stream, d_inputs, 1, ich, w, h, d_weights, k, pad, stride, d_outputs, ch, d_bias, slope);
	if (!d_workspace) { set_error("convolution: this geometry takes the im2col path and needs d_workspace (ich*k*k*Ho*Wo floats)"); return 1; }
	if (im2col_cuda_dev(d_inputs, ich, h, w, k, pad, stride, d_workspace, stream)) return 1;
	const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
	const long long npix = (long long)ho * wo, kk = (long long)ich * k * k;
	if (npix > 0x7fffffffLL || kk > 0x7fffffffLL) { set_error("convolution: problem too large for int dimensions"); return 1; }
	Problem p;
	if (normalise('R', 'N', 'N', ch, (int)npix, (int)kk, 1.f, d_weights, (int)kk, d_workspace, (int)npix, 0.f, d_outputs, (int)npix, &p)) return 1;
	p.bias = d_bias;
	p.slope = slope;
	return run_dev(mode, stream ? static_cast<cudaStream_t>(stream) : g.stream, p);
}

static void conv_host(const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride, float *outputs,
                      int ch, const float *bias, float slope)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	if (ich <= 0 || ch <= 0 || k <= 0 || stride <= 0 || pad < 0 || h + 2 * pad < k || w + 2 * pad < k) { set_error("convolution: bad geometry"); return; }
	const long long ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1, npix = ho * wo, kk = (long long)ich * k * k;
	const size_t in_b = (size_t)ich * h * w * 4, w_b = (size_t)ch * kk * 4, col_b = (size_t)kk * npix * 4, out_b = (size_t)ch * npix * 4, bias_b = bias ? (size_t)ch * 4 : 0;
	const size_t o_in = 0, o_w = align_up(o_in + in_b, 256), o_col = align_up(o_w + w_b, 256), o_out = align_up(o_col + col_b, 256), o_bias = align_up(o_out + out_b, 256);
	if (ensure_arena(o_bias + bias_b + 256)) return;
	float *d_in = reinterpret_cast<float *>(g.arena + o_in), *d_w = reinterpret_cast<float *>(g.arena + o_w);
	float *d_col = reinterpret_cast<float *>(g.arena + o_col), *d_out = reinterpret_cast<float *>(g.arena + o_out);
	float *d_bias = bias ? reinterpret_cast<float *>(g.arena + o_bias) : nullptr;
	cudaError_t e = cudaMemcpyAsync(d_in, inputs, in_b, cudaMemcpyHostToDevice, g.stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, weights, w_b, cudaMemcpyHostToDevice, g.stream);
	if (e == cudaSuccess && bias) e = cudaMemcpyAsync(d_bias, bias, bias_b, cudaMemcpyHostToDevice, g.stream);
	if (e != cudaSuccess) { set_error("convolution H2D failed: %s", cudaGetErrorString(e)); return; }
	if (convolution_cuda_dev(UGEMM_MODE_AUTO, g.stream, d_in, ich, w, h, d_w, k, pad, stride, d_out, ch, d_bias, slope, d_col)) return;
	e = cudaMemcpyAsync(outputs, d_out, out_b, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("convolution failed: %s", cudaGetErrorString(e));
}

void im2col_cuda(const float *im, int channels, int height, int width, int k, int pad, int stride, float *col)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	if (k <= 0 || stride <= 0 || pad < 0 || height + 2 * pad < k || width + 2 * pad < k) { set_error("im2col: bad geometry"); return; }
	const long long ho = (height + 2 * pad - k) / stride + 1, wo = (width + 2 * pad - k) / stride + 1;
	const size_t in_b = (size_t)channels * height * width * 4, col_b = (size_t)channels * k * k * ho * wo * 4;
	const size_t o_col = align_up(in_b, 256);
	if (ensure_arena(o_col + col_b)) return;
	float *d_in = reinterpret_cast<float *>(g.arena), *d_col = reinterpret_cast<float *>(g.arena + o_col);
	cudaError_t e = cudaMemcpyAsync(d_in, im, in_b, cudaMemcpyHostToDevice, g.stream);
	if (e != cudaSuccess) { set_error("im2col H2D failed: %s", cudaGetErrorString(e)); return; }
	if (im2col_cuda_dev(d_in, channels, height, width, k, pad, stride, d_col, g.stream)) return;
	e = cudaMemcpyAsync(col, d_col, col_b, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("im2col failed: %s", cudaGetErrorString(e));
}

void convolution_cuda(const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride, float *outputs, int ch)
{ conv_host(inputs, ich, w, h, weights, k, pad, stride, outputs, ch, nullptr, 1.f); }

void convolution_cuda_LReLU(const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride, float *outputs,
                            int ch, const float *bias)
{ conv_host(inputs, ich, w, h, weights, k, pad, stride, outputs, ch, bias, 0.1f); }

// ---- level 1 / level 2 -------------------------------------------------------------------
int saxpy_cuda_dev(void *stream, int N, float alpha, const float *dx, int incx, float *dy, int incy)
{
	if (ensure_init()) return 1;
	if (N < 0 || incx < 1 || incy < 1) { set_error("saxpy: N must be >= 0 and incx, incy >= 1 (N=%d incx=%d incy=%d)", N, incx, incy); return 1; }
	if (N == 0 || alpha == 0.f) return 0;
	CU_TRY(launch_saxpy(N, alpha, dx, incx, dy, incy, stream ? static_cast<cudaStream_t>(stream) : g.stream, g.sm_count), "saxpy launch");
	g.launches++;
	return 0;
}

void saxpy_cuda(int N, float alpha, const float *x, int incx, float *y, int incy)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	if (N < 0 || incx < 1 || incy < 1) { set_error("saxpy: N must be >= 0 and incx, incy >= 1 (N=%d incx=%d incy=%d)", N, incx, incy); return; }
	if (N == 0 || alpha == 0.f) return;
	const size_t xn = (size_t)(N - 1) * incx + 1, yn = (size_t)(N - 1) * incy + 1;
	const size_t offY = align_up(xn * 4, 256);
	if (ensure_arena(offY + yn * 4)) return;
	float *dx = reinterpret_cast<float *>(g.arena), *dy = reinterpret_cast<float *>(g.arena + offY);
	cudaError_t e = cudaMemcpyAsync(dx, x, xn * 4, cudaMemcpyHostToDevice, g.stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, yn * 4, cudaMemcpyHostToDevice, g.stream);
	if (e != cudaSuccess) { set_error("saxpy H2D failed: %s", cudaGetErrorString(e)); return; }
	if (saxpy_cuda_dev(g.stream, N, alpha, dx, incx, dy, incy)) return;
	// the whole strided span travels both ways, so the gaps between elements come back bit for bit
	e = cudaMemcpyAsync(y, dy, yn * 4, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("saxpy failed: %s", cudaGetErrorString(e));
}

static int gemv_check(char *trans, int M, int N, int lda, int incx, int incy)
{
	*trans = (char)upper(*trans);
	if (*trans != 'N' && *trans != 'T') { set_error("sgemv: trans must be 'N' or 'T'"); return 1; }
	if (M < 0 || N < 0 || incx < 1 || incy < 1) { set_error("sgemv: M, N must be >= 0 and incx, incy >= 1"); return 1; }
	const int need = *trans == 'N' ? M : N;
	if (lda < (need > 1 ? need : 1)) { set_error("sgemv: lda=%d too small for trans='%c' (needs >= %d)", lda, *trans, need); return 1; }
	return 0;
}

int sgemv_cuda_dev(void *stream, char trans, int M, int N, float alpha, const float *dA, int lda, const float *dx, int incx,
                   float beta, float *dy, int incy)
{
	if (ensure_init()) return 1;
	if (gemv_check(&trans, M, N, lda, incx, incy)) return 1;
	if (M == 0) return 0;
	if ((alpha == 0.f || N == 0) && beta == 1.f) return 0;
	// alpha == 0 or N == 0: the sum is empty, y <- beta*y; an N of 0 makes the kernels' loops do exactly that
	CU_TRY(launch_sgemv(trans == 'T', M, alpha == 0.f ? 0 : N, alpha, dA, lda, dx, incx, beta, dy, incy,
	                    stream ? static_cast<cudaStream_t>(stream) : g.stream, g.sm_count), "sgemv launch");
	g.launches++;
	return 0;
}

void sgemv_cuda(char trans, int M, int N, float alpha, const float *A, int lda, const float *x, int incx, float beta, float *y, int incy)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	if (gemv_check(&trans, M, N, lda, incx, incy)) return;
	if (M == 0) return;
	if ((alpha == 0.f || N == 0) && beta == 1.f) return;
	const bool need_ax = alpha != 0.f && N > 0;
	const long long lines = trans == 'N' ? N : M, cols = trans == 'N' ? M : N;
	const size_t a_n = need_ax ? (size_t)((lines - 1) * lda + cols) : 0, xn = need_ax ? (size_t)(N - 1) * incx + 1 : 0, yn = (size_t)(M - 1) * incy + 1;
	const size_t offX = align_up(a_n * 4, 256), offY = align_up(offX + xn * 4, 256);
	if (ensure_arena(offY + yn * 4)) return;
	float *dA = reinterpret_cast<float *>(g.arena), *dx = reinterpret_cast<float *>(g.arena + offX), *dy = reinterpret_cast<float *>(g.arena + offY);
	cudaError_t e = cudaSuccess;
	if (need_ax) {
		e = copy2d(dA, A, lda, lines, cols, cudaMemcpyHostToDevice, g.stream);
		if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, xn * 4, cudaMemcpyHostToDevice, g.stream);
	}
	if (e == cudaSuccess && (beta != 0.f || incy != 1)) e = cudaMemcpyAsync(dy, y, yn * 4, cudaMemcpyHostToDevice, g.stream);
	if (e != cudaSuccess) { set_error("sgemv H2D failed: %s", cudaGetErrorString(e)); return; }
	if (sgemv_cuda_dev(g.stream, trans, M, N, alpha, dA, lda, dx, incx, beta, dy, incy)) return;
	e = cudaMemcpyAsync(y, dy, yn * 4, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("sgemv failed: %s", cudaGetErrorString(e));
}

// ---- DGEMM (check_dgemm.c's path) ----------------------------------------------------------
static int dnormalise(char major, char ta, char tb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                      double beta, double *C, int ldc, DProblem *p)
{
	major = (char)upper(major); ta = (char)upper(ta); tb = (char)upper(tb);
	if (major != 'R' && major != 'C') { set_error("major must be 'R' or 'C' (got 0x%02x)", major); return 1; }
	if ((ta != 'N' && ta != 'T') || (tb != 'N' && tb != 'T')) { set_error("transA/transB must be 'N' or 'T'"); return 1; }
	if (M < 0 || N < 0 || K < 0) { set_error("negative dimension M=%d N=%d K=%d", M, N, K); return 1; }
	if (major == 'C') {   // column-major C = op(A) op(B) is row-major C^T = op(B)^T op(A)^T over the same buffers
		const double *tp = A; A = B; B = tp;
		int ti = lda; lda = ldb; ldb = ti;
		ti = M; M = N; N = ti;
		char tc = ta; ta = tb; tb = tc;
	}
	const int a_cols = ta == 'N' ? K : M, b_cols = tb == 'N' ? N : K;
	if (lda < (a_cols > 1 ? a_cols : 1) || ldb < (b_cols > 1 ? b_cols : 1) || ldc < (N > 1 ? N : 1)) {
		set_error("leading dimension too small (lda=%d ldb=%d ldc=%d for stored widths %d %d %d)", lda, ldb, ldc, a_cols, b_cols, N);
		return 1;
	}
	p->M = M; p->N = N; p->K = K; p->alpha = alpha; p->beta = beta;
	p->A = A; p->lda = lda; p->a_kmajor = (ta == 'N');
	p->B = B; p->ldb = ldb; p->b_kmajor = (tb == 'T');
	p->C = C; p->ldc = ldc;
	return 0;
}

static int drun_dev(cudaStream_t stream, const DProblem &p)
{
	if (p.M == 0 || p.N == 0) return 0;
	if ((p.alpha == 0.0 || p.K == 0) && p.beta == 1.0) return 0;
	if (p.alpha == 0.0 || p.K == 0) CU_TRY(launch_scale_c_f64(p, stream), "scale_c (f64) launch");
	else CU_TRY(launch_k4_dgemm(p, stream), "K4 (DFMA DGEMM) launch");
	g.launches++;
	return 0;
}

int dgemm_cuda_dev(void *stream, char major, char ta, char tb, int M, int N, int K, double alpha, const double *dA, int lda,
                   const double *dB, int ldb, double beta, double *dC, int ldc)
{
	if (ensure_init()) return 1;
	DProblem p;
	if (dnormalise(major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &p)) return 1;
	return drun_dev(stream ? static_cast<cudaStream_t>(stream) : g.stream, p);
}

void dgemm_cuda(char major, char ta, char tb, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                double beta, double *C, int ldc)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (ensure_init()) return;
	DProblem p;
	if (dnormalise(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &p)) return;
	if (p.M == 0 || p.N == 0) return;
	if ((p.alpha == 0.0 || p.K == 0) && p.beta == 1.0) return;
	const bool need_ab = !(p.alpha == 0.0 || p.K == 0);
	const long long a_lines = p.a_kmajor ? p.M : p.K, a_cols = p.a_kmajor ? p.K : p.M;
	const long long b_lines = p.b_kmajor ? p.N : p.K, b_cols = p.b_kmajor ? p.K : p.N;
	const size_t a_bytes = need_ab ? (size_t)((a_lines - 1) * p.lda + a_cols) * 8 : 0;
	const size_t b_bytes = need_ab ? (size_t)((b_lines - 1) * p.ldb + b_cols) * 8 : 0;
	const size_t c_bytes = (size_t)((long long)(p.M - 1) * p.ldc + p.N) * 8;
	const size_t offB = align_up(a_bytes, 256), offC = align_up(offB + b_bytes, 256);
	if (ensure_arena(offC + c_bytes)) return;
	double *dA = reinterpret_cast<double *>(g.arena), *dB = reinterpret_cast<double *>(g.arena + offB), *dC = reinterpret_cast<double *>(g.arena + offC);
	cudaError_t e = cudaSuccess;
	auto up2d = [&](double *d, const double *h, long long ld, long long lines, long long cols) {
		if (e != cudaSuccess || lines <= 0 || cols <= 0) return;
		if (ld == cols) e = cudaMemcpyAsync(d, h, (size_t)lines * cols * 8, cudaMemcpyHostToDevice, g.stream);
		else e = cudaMemcpy2DAsync(d, (size_t)ld * 8, h, (size_t)ld * 8, (size_t)cols * 8, (size_t)lines, cudaMemcpyHostToDevice, g.stream);
	};
	if (need_ab) { up2d(dA, p.A, p.lda, a_lines, a_cols); up2d(dB, p.B, p.ldb, b_lines, b_cols); }
	if (p.beta != 0.0) up2d(dC, p.C, p.ldc, p.M, p.N);
	if (e != cudaSuccess) { set_error("dgemm H2D copy failed: %s", cudaGetErrorString(e)); return; }
	DProblem d = p;
	d.A = dA; d.B = dB; d.C = dC;
	if (drun_dev(g.stream, d)) return;
	if (p.ldc == p.N) e = cudaMemcpyAsync(p.C, dC, (size_t)p.M * p.N * 8, cudaMemcpyDeviceToHost, g.stream);
	else e = cudaMemcpy2DAsync(p.C, (size_t)p.ldc * 8, dC, (size_t)p.ldc * 8, (size_t)p.N * 8, (size_t)p.M, cudaMemcpyDeviceToHost, g.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
	if (e != cudaSuccess) set_error("dgemm failed: %s", cudaGetErrorString(e));
}

int dgemm_cuda_time_dev(int iters, int warmup, char major, char ta, char tb, int M, int N, int K, double alpha, const double *dA, int lda,
                        const double *dB, int ldb, double beta, double *dC, int ldc, float *ms_avg, float *ms_min)
{
	if (ensure_init()) return 1;
	if (iters < 1) iters = 1;
	if (iters > 1024) iters = 1024;
	DProblem p;
	if (dnormalise(major, ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc, &p)) return 1;
	for (int i = 0; i < warmup; i++)
		if (drun_dev(g.stream, p)) return 1;
	CU_TRY(cudaStreamSynchronize(g.stream), "warm-up sync");
	cudaEvent_t ev[2];
	float total = 0.f, best = 1e30f;
	CU_TRY(cudaEventCreate(&ev[0]), "cudaEventCreate");
	CU_TRY(cudaEventCreate(&ev[1]), "cudaEventCreate");
	int rc = 0;
	for (int i = 0; i < iters && !rc; i++) {
		cudaEventRecord(ev[0], g.stream);
		rc = drun_dev(g.stream, p);
		cudaEventRecord(ev[1], g.stream);
		if (!rc && cudaEventSynchronize(ev[1]) != cudaSuccess) { set_error("timed dgemm launch failed"); rc = 1; }
		if (!rc) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[0], ev[1]); total += ms; if (ms < best) best = ms; }
	}
	cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
	if (!rc) { if (ms_avg) *ms_avg = total / iters; if (ms_min) *ms_min = best; }
	return rc;
}

int ugemm_cuda_probe_tf32(const float *A, const float *B, float *D, int ksteps)
{
	if (ensure_init()) return 1;
	if (ksteps < 1 || ksteps > 4) { set_error("probe: ksteps must be 1..4"); return 1; }
	const size_t abytes = (size_t)128 * 8 * ksteps * 4, bbytes = (size_t)16 * 8 * ksteps * 4, dbytes = 128 * 16 * 4;
	float *dA = nullptr, *dB = nullptr, *dD = nullptr;
	CU_TRY(cudaMalloc(&dA, abytes), "probe cudaMalloc");
	CU_TRY(cudaMalloc(&dB, bbytes), "probe cudaMalloc");
	CU_TRY(cudaMalloc(&dD, dbytes), "probe cudaMalloc");
	int rc = 1;
	do {
		if (cudaMemcpyAsync(dA, A, abytes, cudaMemcpyHostToDevice, g.stream) != cudaSuccess) break;
		if (cudaMemcpyAsync(dB, B, bbytes, cudaMemcpyHostToDevice, g.stream) != cudaSuccess) break;
		if (launch_probe_tf32(dA, dB, dD, ksteps, g.stream) != cudaSuccess) break;
		g.launches++;
		if (cudaMemcpyAsync(D, dD, dbytes, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess) break;
		cudaError_t e = cudaStreamSynchronize(g.stream);
		if (e != cudaSuccess) { set_error("probe failed: %s", cudaGetErrorString(e)); break; }
		rc = 0;
	} while (0);
	if (rc && !g_has_err) set_error("probe failed: %s", cudaGetErrorString(cudaGetLastError()));
	cudaFree(dA); cudaFree(dB); cudaFree(dD);
	return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// Single-process multi-GPU SGEMM (SURVEY §8 e): the C host program's way to the sharded path (ugemm_b200/dist.py is the
// one-process-per-GPU twin used by bench.py).  C is cut into a pr x pc grid of blocks; device d = i*pc + j owns block
// (i, j) and needs A row-panel i and B column-panel j.  K is not split across devices: one exchange step, no reduction.
//
// Distribution is a pipelined relay over NVLink on the copy engines (no SM is taken from the GEMM): K is cut into slabs;
// slab t of A panel i is pulled by device (i, 0) from the caller's buffer and then by (i, 1) from (i, 0), (i, 2) from
// (i, 1) ...; B panels relay down the grid columns the same way -- so the root's egress is one copy of A and one of B,
// every device forwards at most its own two panels, and slab t+1 travels while slab t is multiplied (beta = 1 after the
// first slab).  The local product is run_dev(), i.e. the same rule-based K1/K2 launch as sgemm_cuda_dev.
// A, B, C may be pinned/pageable host memory or device memory of any GPU (unified addressing; copies use
// cudaMemcpyDefault).  Blocking, like every host-pointer entry point.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int MG_MAX_DEV = 16, MG_MAX_SLABS = 64;
struct MgDev {
	cudaStream_t copy_a = nullptr, copy_b = nullptr, comp = nullptr;
	cudaEvent_t landed_a[MG_MAX_SLABS] = {nullptr}, landed_b[MG_MAX_SLABS] = {nullptr};
	cudaEvent_t e_start = nullptr, e_a_done = nullptr, e_b_done = nullptr, e_g0 = nullptr, e_g1 = nullptr, e_end = nullptr;
	char *arena = nullptr;
	size_t arena_bytes = 0;
};
struct MgState { int n = 0; MgDev d[MG_MAX_DEV]; } mg;

int mg_arena(int dev, size_t bytes)
{
	MgDev &D = mg.d[dev];
	if (bytes <= D.arena_bytes) return 0;
	if (D.arena) cudaFree(D.arena);
	D.arena = nullptr; D.arena_bytes = 0;
	const size_t want = bytes + (bytes >> 4) + (1u << 20);
	CU_TRY(cudaMalloc(&D.arena, want), "multi-GPU arena cudaMalloc");
	D.arena_bytes = want;
	return 0;
}

// The partition of one call (row-major view): block sizes (multiples of 4 so local leading dimensions stay TMA-eligible) and
// K slabs (multiples of 32).  8192-wide slabs: every slab after the first is a beta = 1 pass over the C block (+2..6 % per pass
// at 4096), while the exposed part of the distribution is only the first slab's transfer.  overlap > 1 forces the slab pipeline
// on a 1 x 1 grid too (tests).
void mg_plan(int M, int N, int K, int pr, int pc, int overlap, int *mb, int *nb, int *slabs, int *kw)
{
	*mb = ((M + pr - 1) / pr + 3) / 4 * 4;
	*nb = ((N + pc - 1) / pc + 3) / 4 * 4;
	int L = 1, w = K;
	if (K > 0 && ((overlap && pr * pc > 1) || overlap > 1)) {
		L = (K + 8191) / 8192;
		if (L > MG_MAX_SLABS) L = MG_MAX_SLABS;
		w = ((K + L - 1) / L + 31) / 32 * 32;
		L = (K + w - 1) / w;
	}
	*slabs = L; *kw = w;
}

// 2-D copy between any two pointers of the unified address space; lines of `cols` floats, pitches in floats
cudaError_t mg_copy(float *dst, long long dld, const float *src, long long sld, long long lines, long long cols, cudaStream_t st)
{
	if (lines <= 0 || cols <= 0) return cudaSuccess;
	if (dld == cols && sld == cols) return cudaMemcpyAsync(dst, src, (size_t)lines * cols * 4, cudaMemcpyDefault, st);
	return cudaMemcpy2DAsync(dst, (size_t)dld * 4, src, (size_t)sld * 4, (size_t)cols * 4, (size_t)lines, cudaMemcpyDefault, st);
}

} // namespace

int sgemm_cuda_mgpu_init(int ngpus)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (mg.n) return ngpus == mg.n ? 0 : (set_error("sgemm_cuda_mgpu_init: already initialised with %d GPUs", mg.n), 1);
	if (ensure_init()) return 1;
	if (g.device != 0) { set_error("sgemm_cuda_mgpu_init: the backend must be initialised on device 0 (it is on %d)", g.device); return 1; }
	int count = 0;
	CU_TRY(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
	if (ngpus < 1 || ngpus > count || ngpus > MG_MAX_DEV) { set_error("sgemm_cuda_mgpu_init: %d GPUs requested, %d visible (max %d)", ngpus, count, MG_MAX_DEV); return 1; }
	for (int d = 0; d < ngpus; d++) {
		cudaDeviceProp prop;
		CU_TRY(cudaGetDeviceProperties(&prop, d), "cudaGetDeviceProperties");
		if (prop.major != 10) { set_error("device %d (%s) is sm_%d%d; this backend is built for sm_100a only", d, prop.name, prop.major, prop.minor); return 1; }
		CU_TRY(cudaSetDevice(d), "cudaSetDevice");
		MgDev &D = mg.d[d];
		CU_TRY(cudaStreamCreateWithFlags(&D.copy_a, cudaStreamNonBlocking), "cudaStreamCreate");
		CU_TRY(cudaStreamCreateWithFlags(&D.copy_b, cudaStreamNonBlocking), "cudaStreamCreate");
		CU_TRY(cudaStreamCreateWithFlags(&D.comp, cudaStreamNonBlocking), "cudaStreamCreate");
		for (int t = 0; t < MG_MAX_SLABS; t++) {
			CU_TRY(cudaEventCreateWithFlags(&D.landed_a[t], cudaEventDisableTiming), "cudaEventCreate");
			CU_TRY(cudaEventCreateWithFlags(&D.landed_b[t], cudaEventDisableTiming), "cudaEventCreate");
		}
		cudaEvent_t *timed[] = {&D.e_start, &D.e_a_done, &D.e_b_done, &D.e_g0, &D.e_g1, &D.e_end};
		for (cudaEvent_t *e : timed) CU_TRY(cudaEventCreate(e), "cudaEventCreate");
		for (int q = 0; q < ngpus; q++) {
			if (q == d) continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, d, q);
			if (!can) { set_error("sgemm_cuda_mgpu_init: device %d cannot access device %d (no NVLink/PCIe peer path)", d, q); cudaSetDevice(g.device); return 1; }
			cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", d, q, cudaGetErrorString(e)); cudaSetDevice(g.device); return 1; }
			cudaGetLastError();
		}
		cudaMemPool_t pool;   // the repack path's stream-ordered scratch: keep it cached, like on device 0
		if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) { unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
		cudaGetLastError();
	}
	CU_TRY(cudaSetDevice(g.device), "cudaSetDevice");
	mg.n = ngpus;
	return 0;
}

void sgemm_cuda_mgpu_finish(void)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	for (int d = 0; d < mg.n; d++) {
		MgDev &D = mg.d[d];
		cudaSetDevice(d);
		cudaDeviceSynchronize();
		if (D.arena) cudaFree(D.arena);
		cudaStreamDestroy(D.copy_a); cudaStreamDestroy(D.copy_b); cudaStreamDestroy(D.comp);
		for (int t = 0; t < MG_MAX_SLABS; t++) { cudaEventDestroy(D.landed_a[t]); cudaEventDestroy(D.landed_b[t]); }
		cudaEventDestroy(D.e_start); cudaEventDestroy(D.e_a_done); cudaEventDestroy(D.e_b_done);
		cudaEventDestroy(D.e_g0); cudaEventDestroy(D.e_g1); cudaEventDestroy(D.e_end);
		D = MgDev();
	}
	if (mg.n) cudaSetDevice(g.device);
	mg.n = 0;
}

int sgemm_cuda_mgpu_count(void) { return mg.n; }

int sgemm_cuda_mgpu_plan(int M, int N, int K, int pr, int pc, int overlap, int *block_rows, int *block_cols, int *k_slabs, int *slab_width)
{
	if (M < 0 || N < 0 || K < 0 || pr < 1 || pc < 1 || (long long)pr * pc > MG_MAX_DEV) return 1;
	int mb, nb, L, kw;
	mg_plan(M, N, K, pr, pc, overlap, &mb, &nb, &L, &kw);
	if (block_rows) *block_rows = mb;
	if (block_cols) *block_cols = nb;
	if (k_slabs) *k_slabs = L;
	if (slab_width) *slab_width = kw;
	return 0;
}

int ugemm_cuda_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int sgemm_cuda_mgpu(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
                    const float *B, int ldb, float beta, float *C, int ldc, int pr, int pc, int overlap, float *timings_ms)
{
	std::lock_guard<std::recursive_mutex> lk(g_mu);
	if (timings_ms) for (int i = 0; i < 5; i++) timings_ms[i] = 0.f;
	if (!mg.n) { set_error("sgemm_cuda_mgpu: call sgemm_cuda_mgpu_init first"); return 1; }
	if (pr < 1 || pc < 1 || (long long)pr * pc > mg.n) { set_error("sgemm_cuda_mgpu: grid %d x %d needs %lld GPUs, %d initialised", pr, pc, (long long)pr * pc, mg.n); return 1; }
	Problem p;
	if (normalise(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &p)) return 1;
	if (major == 'C' || major == 'c') { int t = pr; pr = pc; pc = t; }   // the grid follows the row-major view of C
	if (p.M == 0 || p.N == 0) return 0;
	const bool scale_only = p.alpha == 0.f || p.K == 0;
	if (scale_only && p.beta == 1.f) return 0;

	int mb, nb, L, kw;
	mg_plan(p.M, p.N, p.K, pr, pc, overlap, &mb, &nb, &L, &kw);
	if (scale_only) L = 0;

	// which device (if any) holds a caller pointer: a block whose source already lives on the device that needs it is used
	// in place (no local copy) -- with the operands on GPU 0, GPU 0 starts multiplying at once and only serves its peers
	auto home_of = [](const void *ptr) {
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return -1; }
		return at.type == cudaMemoryTypeDevice ? at.device : -1;
	};
	const int home_a = scale_only ? -1 : home_of(p.A), home_b = scale_only ? -1 : home_of(p.B), home_c = home_of(p.C);
	struct Loc { int mi, nj; float *A, *B, *C; long long lda, ldb, ldc; bool on, a_inplace, b_inplace, c_inplace; };
	Loc loc[MG_MAX_DEV];
	const int nd = pr * pc;
	int rc = 0;
	for (int d = 0; d < nd && !rc; d++) {
		const int i = d / pc, j = d % pc;
		Loc &l = loc[d];
		l.mi = p.M - i * mb < mb ? p.M - i * mb : mb;
		l.nj = p.N - j * nb < nb ? p.N - j * nb : nb;
		l.on = l.mi > 0 && l.nj > 0;
		if (!l.on) continue;
		// local panels keep the caller's orientation: K-major = lines of K floats, MN-major = K lines of mi (nj) floats
		l.lda = p.a_kmajor ? (p.K + 3) / 4 * 4 : mb;
		l.ldb = p.b_kmajor ? (p.K + 3) / 4 * 4 : nb;
		l.ldc = nb;
		l.a_inplace = j == 0 && home_a == d;
		l.b_inplace = i == 0 && home_b == d;
		l.c_inplace = home_c == d;
		const size_t a_bytes = l.a_inplace ? 0 : align_up_sz((size_t)(p.a_kmajor ? l.mi : p.K) * l.lda * 4, 256);
		const size_t b_bytes = l.b_inplace ? 0 : align_up_sz((size_t)(p.b_kmajor ? l.nj : p.K) * l.ldb * 4, 256);
		const size_t c_bytes = l.c_inplace ? 0 : align_up_sz((size_t)l.mi * l.ldc * 4, 256);
		if (cudaSetDevice(d) != cudaSuccess || mg_arena(d, a_bytes + b_bytes + c_bytes + 256)) { rc = 1; break; }
		l.A = reinterpret_cast<float *>(mg.d[d].arena);
		l.B = reinterpret_cast<float *>(mg.d[d].arena + a_bytes);
		l.C = reinterpret_cast<float *>(mg.d[d].arena + a_bytes + b_bytes);
		if (l.a_inplace) { l.A = const_cast<float *>(p.A) + (p.a_kmajor ? (long long)i * mb * p.lda : (long long)i * mb); l.lda = p.lda; }
		if (l.b_inplace) { l.B = const_cast<float *>(p.B) + (p.b_kmajor ? (long long)j * nb * p.ldb : (long long)j * nb); l.ldb = p.ldb; }
		if (l.c_inplace) { l.C = p.C + (long long)i * mb * p.ldc + (long long)j * nb; l.ldc = p.ldc; }
	}
	cudaError_t e = cudaSuccess;
	auto ok = [&]() { return rc == 0 && e == cudaSuccess; };
	timespec ts0, ts1;
	clock_gettime(CLOCK_MONOTONIC, &ts0);

	// start marks + C blocks in (only when they are read)
	for (int d = 0; d < nd && ok(); d++) {
		if (!loc[d].on) continue;
		const int i = d / pc, j = d % pc;
		MgDev &D = mg.d[d];
		e = cudaSetDevice(d);
		if (e == cudaSuccess) e = cudaEventRecord(D.e_start, D.comp);
		if (e == cudaSuccess && p.beta != 0.f && !loc[d].c_inplace)
			e = mg_copy(loc[d].C, loc[d].ldc, p.C + (long long)i * mb * p.ldc + (long long)j * nb, p.ldc, loc[d].mi, loc[d].nj, D.comp);
	}
	// panel relay, issued slab-major so that every device sees slab 0 first
	for (int t = 0; t < L && ok(); t++) {
		const int k0 = t * kw, kt = p.K - k0 < kw ? p.K - k0 : kw;
		for (int d = 0; d < nd && ok(); d++) {
			if (!loc[d].on) continue;
			const int i = d / pc, j = d % pc;
			MgDev &D = mg.d[d];
			const Loc &l = loc[d];
			e = cudaSetDevice(d);
			if (e != cudaSuccess) break;
			{   // A panel i, slab t: from the caller's buffer (j == 0) or from the left neighbour's copy
				const long long lines = p.a_kmajor ? l.mi : kt, cols = p.a_kmajor ? kt : l.mi;
				float *dst = l.A + (p.a_kmajor ? (long long)k0 : (long long)k0 * l.lda);
				if (l.a_inplace) {
					// already here: nothing to move, the slab has "landed"
				} else if (j == 0) {
					const float *src = p.A + (p.a_kmajor ? (long long)i * mb * p.lda + k0 : (long long)k0 * p.lda + (long long)i * mb);
					e = mg_copy(dst, l.lda, src, p.lda, lines, cols, D.copy_a);
				} else {
					const Loc &s = loc[d - 1];
					e = cudaStreamWaitEvent(D.copy_a, mg.d[d - 1].landed_a[t], 0);
					if (e == cudaSuccess) e = mg_copy(dst, l.lda, s.A + (p.a_kmajor ? (long long)k0 : (long long)k0 * s.lda), s.lda, lines, cols, D.copy_a);
				}
				if (e == cudaSuccess) e = cudaEventRecord(D.landed_a[t], D.copy_a);
				if (e == cudaSuccess && t == L - 1) e = cudaEventRecord(D.e_a_done, D.copy_a);
			}
			if (e != cudaSuccess) break;
			{   // B panel j, slab t: from the caller's buffer (i == 0) or from the upper neighbour's copy
				const long long lines = p.b_kmajor ? l.nj : kt, cols = p.b_kmajor ? kt : l.nj;
				float *dst = l.B + (p.b_kmajor ? (long long)k0 : (long long)k0 * l.ldb);
				if (l.b_inplace) {
				} else if (i == 0) {
					const float *src = p.B + (p.b_kmajor ? (long long)j * nb * p.ldb + k0 : (long long)k0 * p.ldb + (long long)j * nb);
					e = mg_copy(dst, l.ldb, src, p.ldb, lines, cols, D.copy_b);
				} else {
					const Loc &s = loc[d - pc];
					e = cudaStreamWaitEvent(D.copy_b, mg.d[d - pc].landed_b[t], 0);
					if (e == cudaSuccess) e = mg_copy(dst, l.ldb, s.B + (p.b_kmajor ? (long long)k0 : (long long)k0 * s.ldb), s.ldb, lines, cols, D.copy_b);
				}
				if (e == cudaSuccess) e = cudaEventRecord(D.landed_b[t], D.copy_b);
				if (e == cudaSuccess && t == L - 1) e = cudaEventRecord(D.e_b_done, D.copy_b);
			}
		}
	}
	// local products: slab t starts when ITS two copies have landed
	for (int d = 0; d < nd && ok(); d++) {
		if (!loc[d].on) continue;
		MgDev &D = mg.d[d];
		const Loc &l = loc[d];
		e = cudaSetDevice(d);
		if (e != cudaSuccess) break;
		Problem q = p;
		q.M = l.mi; q.N = l.nj; q.lda = l.lda; q.ldb = l.ldb; q.C = l.C; q.ldc = l.ldc;
		if (scale_only) {
			e = cudaEventRecord(D.e_g0, D.comp);
			q.A = l.A; q.B = l.B;
			if (e == cudaSuccess && run_dev(UGEMM_MODE_AUTO, D.comp, q)) rc = 1;
		}
		for (int t = 0; t < L && ok(); t++) {
			const int k0 = t * kw, kt = p.K - k0 < kw ? p.K - k0 : kw;
			e = cudaStreamWaitEvent(D.comp, D.landed_a[t], 0);
			if (e == cudaSuccess) e = cudaStreamWaitEvent(D.comp, D.landed_b[t], 0);
			if (e == cudaSuccess && t == 0) e = cudaEventRecord(D.e_g0, D.comp);
			if (e != cudaSuccess) break;
			q.K = kt;
			q.A = l.A + (p.a_kmajor ? (long long)k0 : (long long)k0 * l.lda);
			q.B = l.B + (p.b_kmajor ? (long long)k0 : (long long)k0 * l.ldb);
			q.beta = t == 0 ? p.beta : 1.f;
			if (run_dev(UGEMM_MODE_AUTO, D.comp, q)) rc = 1;
		}
		if (ok()) e = cudaEventRecord(D.e_g1, D.comp);
		// the finished block goes back into the caller's C: only its mi x nj region, ld padding is never written
		const int i = d / pc, j = d % pc;
		if (ok() && !l.c_inplace) e = mg_copy(p.C + (long long)i * mb * p.ldc + (long long)j * nb, p.ldc, l.C, l.ldc, l.mi, l.nj, D.comp);
		if (ok()) e = cudaEventRecord(D.e_end, D.comp);
	}
	// drain every device (also after an error, so nothing is left in flight on the arenas)
	float t_span = 0.f, t_dist = 0.f, t_gemm = 0.f, t_done = 0.f;
	for (int d = 0; d < nd; d++) {
		if (!loc[d].on) continue;
		MgDev &D = mg.d[d];
		cudaSetDevice(d);
		cudaError_t es = cudaStreamSynchronize(D.copy_a);
		cudaError_t e2 = cudaStreamSynchronize(D.copy_b); if (es == cudaSuccess) es = e2;
		e2 = cudaStreamSynchronize(D.comp); if (es == cudaSuccess) es = e2;
		if (es != cudaSuccess && e == cudaSuccess) e = es;
		if (ok()) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, D.e_start, D.e_end) == cudaSuccess && ms > t_span) t_span = ms;
			if (L > 0) {
				if (cudaEventElapsedTime(&ms, D.e_start, D.e_a_done) == cudaSuccess && ms > t_dist) t_dist = ms;
				if (cudaEventElapsedTime(&ms, D.e_start, D.e_b_done) == cudaSuccess && ms > t_dist) t_dist = ms;
			}
			if (cudaEventElapsedTime(&ms, D.e_g0, D.e_g1) == cudaSuccess && ms > t_gemm) t_gemm = ms;
			if (cudaEventElapsedTime(&ms, D.e_start, D.e_g1) == cudaSuccess && ms > t_done) t_done = ms;
			cudaGetLastError();
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &ts1);
	cudaSetDevice(g.device);
	if (e != cudaSuccess && !g_has_err) {
		const unsigned *dg = k1_diag_host();
		if (dg && dg[0]) set_error("sgemm_cuda_mgpu failed: %s (K1 watchdog code %u, block %u, thread %u)", cudaGetErrorString(e), dg[0], dg[1], dg[2]);
		else set_error("sgemm_cuda_mgpu failed: %s", cudaGetErrorString(e));
	}
	if (!ok()) return 1;
	if (timings_ms) {
		timings_ms[0] = (float)((ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);   // host wall clock, issue included
		timings_ms[1] = t_span;    // max over devices: start mark -> C block written back
		timings_ms[2] = t_dist;    // max over devices: start mark -> last panel slab landed
		timings_ms[3] = t_gemm;    // max over devices: first product start -> last product end (overlap=0: compute only)
		timings_ms[4] = t_done;    // max over devices: start mark -> last product end (distribution in, C write-back out)
	}
	return 0;
}

int sgemm_cuda_mgpu_run(char major, char ta, char tb, int M, int N, int K, float alpha, const float *A, int lda,
                        const float *B, int ldb, float beta, float *C, int ldc, int pr, int pc, int overlap, float *timings_ms)
{ return sgemm_cuda_mgpu(major, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, pr, pc, overlap, timings_ms); }

} // extern "C"
