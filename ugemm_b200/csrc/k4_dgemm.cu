// k4_dgemm.cu -- K4: DGEMM for sm_100a on the FP64 tensor path (the double-precision sibling of K2).
//
// Replaces the reference's dgemm_cpu (ugemm.h:162-285), _dgemm_c (gemm_cpu.h instantiated with real = double,
// ugemm.h:29-33) and dgemm_avx (dgemm_avx.h:844-925) -- the three rows of check_dgemm.c:255-258 -- for device data.
// Same semantics as the SGEMM path: C = alpha * op(A) op(B) + beta * C, row- or column-major, N/T, any ld.
//
// Shape of the kernel: 128 x 64 C tile per CTA, BK = 8, 256 threads.  The products run on the FP64 tensor path (DMMA,
// mma.sync.m8n8k4.f64 -- there is no tcgen05 kind for fp64): each warp owns a 32 x 32 block of the tile as 4 x 4 MMA tiles
// (one double of A and of B per lane and tile, two accumulators per lane and tile).  -DUGEMM_K4_DMMA=0 builds the plain DFMA
// variant (8 x 4 register tile per thread, halves per dimension so the shared-memory reads are 128-bit).  Global loads are
// 128-bit when the operand allows it (base 16 B aligned, ld even); interior tiles take the same unguarded steady-state loop
// as K2; transposes are folded into the global->shared staging.  Bound: the FP64 pipe, 148 SMs x 64 lanes x 2 x clock
// = 37.2 TFLOP/s at 1965 MHz.  Algorithmic cost 2*M*N*K flop, 8*(MK + KN + MN(1+[beta != 0])) bytes.
#include "common.cuh"

namespace ugemm {

namespace {

constexpr int D_BK = 8;
constexpr int D_THREADS = 256;
constexpr int D_PAD = 4;      // row pitch = 4 mod 16 doubles: the 4 k-rows x 8 lines of an MMA fragment load hit 16 distinct 8-byte banks
#ifndef UGEMM_K4_DMMA
#define UGEMM_K4_DMMA 1
#endif
constexpr int D_BM = 128, D_BN = 64, D_TM = 8, D_TN = 4;

__device__ __forceinline__ double2 load_pair(const double *__restrict__ line, long long c, long long cmax, bool line_ok, bool vec)
{
	double2 v = make_double2(0.0, 0.0);
	if (!line_ok) return v;
	if (vec && c + 1 < cmax) return __ldg(reinterpret_cast<const double2 *>(line + c));
	if (c + 0 < cmax) v.x = __ldg(line + c + 0);
	if (c + 1 < cmax) v.y = __ldg(line + c + 1);
	return v;
}

// Operand staging, as in K2 (k2_simt.cu: Stager) with pairs of doubles instead of quads of floats.
template <int BMN, bool KCONTIG>
struct DStager {
	static constexpr int NP = BMN * D_BK / 2;
	static constexpr int PAIRS = NP / D_THREADS;
	static_assert(NP % D_THREADS == 0, "tile must be a whole number of pairs per thread");
	double2 r[PAIRS];
	const double *fp;
	long long fstep;

	__device__ __forceinline__ void load(const double *__restrict__ base, long long ld, long long mn0, long long mn_max,
	                                     long long k0, long long k_max, bool vec, int tid)
	{
#pragma unroll
		for (int i = 0; i < PAIRS; i++) {
			const int f = tid + i * D_THREADS;
			if (KCONTIG) {
				const int line = f / (D_BK / 2), kp = (f % (D_BK / 2)) * 2;
				const long long mn = mn0 + line;
				r[i] = load_pair(base + mn * ld, k0 + kp, k_max, mn < mn_max, vec);
			} else {
				const int k = f / (BMN / 2), q = (f % (BMN / 2)) * 2;
				const long long kk = k0 + k;
				r[i] = load_pair(base + kk * ld, mn0 + q, mn_max, kk < k_max, vec);
			}
		}
	}
	__device__ __forceinline__ void fast_init(const double *__restrict__ base, long long ld, long long mn0, int tid)
	{
		if (KCONTIG) { fp = base + (mn0 + tid / (D_BK / 2)) * ld + (tid % (D_BK / 2)) * 2; fstep = (D_THREADS / (D_BK / 2)) * ld; }
		else         { fp = base + (long long)(tid / (BMN / 2)) * ld + mn0 + (tid % (BMN / 2)) * 2; fstep = (D_THREADS / (BMN / 2)) * ld; }
	}
	__device__ __forceinline__ void fast_load(long long ld)
	{
		fp += KCONTIG ? (long long)D_BK : D_BK * ld;
#pragma unroll
		for (int i = 0; i < PAIRS; i++) r[i] = __ldg(reinterpret_cast<const double2 *>(fp + i * fstep));
	}
	__device__ __forceinline__ void store(double (*s)[BMN + D_PAD], int tid) const
	{
#pragma unroll
		for (int i = 0; i < PAIRS; i++) {
			const int f = tid + i * D_THREADS;
			if (KCONTIG) {
				const int line = f / (D_BK / 2), kp = (f % (D_BK / 2)) * 2;
				s[kp + 0][line] = r[i].x; s[kp + 1][line] = r[i].y;
			} else {
				const int k = f / (BMN / 2), q = (f % (BMN / 2)) * 2;
				*reinterpret_cast<double2 *>(&s[k][q]) = r[i];
			}
		}
	}
};

template <bool AK, bool BKM>
__global__ void __launch_bounds__(D_THREADS, 2)
k4_dgemm_kernel(DProblem p, const int tiles_m, const int tiles_n, const bool vecA, const bool vecB, const bool vecC)
{
	constexpr int BM = D_BM, BN = D_BN, TM = D_TM, TN = D_TN, HM = TM / 2, HN = TN / 2;
	static_assert((BM / TM) * (BN / TN) == D_THREADS, "thread tile must cover the CTA tile");
	__shared__ __align__(16) double As[2][D_BK][BM + D_PAD];
	__shared__ __align__(16) double Bs[2][D_BK][BN + D_PAD];

	const int tid = threadIdx.x;
	const int tx = tid % (BN / TN), ty = tid / (BN / TN);

	// grouped tile order: 16 consecutive m-tiles share the same n-tile sweep (keeps A/B panels in L2)
	constexpr int GROUP = 16;
	const long long tile = blockIdx.x;
	const long long per_group = (long long)GROUP * tiles_n;
	const int group = (int)(tile / per_group);
	const int first_m = group * GROUP;
	const int gsize = min(tiles_m - first_m, GROUP);
	const int tm = first_m + (int)((tile % per_group) % gsize);
	const int tn = (int)((tile % per_group) / gsize);
	const long long m0 = (long long)tm * BM, n0 = (long long)tn * BN;

#if UGEMM_K4_DMMA
	// warp w: rows (w / 2) * 32 .. +32, columns (w % 2) * 32 .. +32 of the CTA tile; lane = 4 * g + q.
	// m8n8k4 fragments: A[row g][k q], B[k q][col g], C[row g][cols 2q, 2q+1]
	const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
	const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
	double acc[4][4][2];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#else
	double acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; i++)
#pragma unroll
		for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
#endif

	DStager<BM, AK> sa;
	DStager<BN, BKM> sb;
	const int ktiles = (p.K + D_BK - 1) / D_BK;
	const bool interior = vecA && vecB && m0 + BM <= p.M && n0 + BN <= p.N;
	const int fast_tiles = interior ? p.K / D_BK : 0;
	sa.fast_init(p.A, p.lda, m0, tid);
	sb.fast_init(p.B, p.ldb, n0, tid);
	sa.load(p.A, p.lda, m0, p.M, 0, p.K, vecA, tid);
	sb.load(p.B, p.ldb, n0, p.N, 0, p.K, vecB, tid);
	sa.store(As[0], tid);
	sb.store(Bs[0], tid);
	__syncthreads();

	auto multiply = [&](int cur) {
#if UGEMM_K4_DMMA
#pragma unroll
		for (int k4 = 0; k4 < D_BK; k4 += 4) {
			double a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) a[i] = As[cur][k4 + q][wm + i * 8 + g];
#pragma unroll
			for (int j = 0; j < 4; j++) b[j] = Bs[cur][k4 + q][wn + j * 8 + g];
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++)
					asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
					             : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
		}
		return;
#endif
#pragma unroll
		for (int kk = 0; kk < D_BK; kk++) {
#if !UGEMM_K4_DMMA
			double a[TM], b[TN];
#pragma unroll
			for (int i = 0; i < HM; i++) {
				a[i] = As[cur][kk][ty * HM + i];
				a[HM + i] = As[cur][kk][BM / 2 + ty * HM + i];
			}
#pragma unroll
			for (int j = 0; j < HN; j++) {
				b[j] = Bs[cur][kk][tx * HN + j];
				b[HN + j] = Bs[cur][kk][BN / 2 + tx * HN + j];
			}
#pragma unroll
			for (int i = 0; i < TM; i++)
#pragma unroll
				for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
#endif
		}
	};

	int t = 0;
	for (; t + 1 < fast_tiles; t++) {
		const int cur = t & 1;
		sa.fast_load(p.lda);
		sb.fast_load(p.ldb);
		multiply(cur);
		sa.store(As[cur ^ 1], tid);
		sb.store(Bs[cur ^ 1], tid);
		__syncthreads();
	}
	for (; t < ktiles; t++) {
		const int cur = t & 1;
		if (t + 1 < ktiles) {
			sa.load(p.A, p.lda, m0, p.M, (long long)(t + 1) * D_BK, p.K, vecA, tid);
			sb.load(p.B, p.ldb, n0, p.N, (long long)(t + 1) * D_BK, p.K, vecB, tid);
		}
		multiply(cur);
		if (t + 1 < ktiles) {
			sa.store(As[cur ^ 1], tid);
			sb.store(Bs[cur ^ 1], tid);
		}
		__syncthreads();
	}

	// fused epilogue: C = alpha*acc + beta*C (C never read when beta == 0), ld padding never touched
	const double alpha = p.alpha, beta = p.beta;
#if UGEMM_K4_DMMA
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const long long m = m0 + wm + i * 8 + g;
		if (m >= p.M) continue;
		double *crow = p.C + m * p.ldc;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const long long n = n0 + wn + j * 8 + 2 * q;
			if (vecC && n + 1 < p.N) {
				double2 *cp = reinterpret_cast<double2 *>(crow + n);
				double2 o;
				if (beta != 0.0) {
					const double2 c = *cp;
					o.x = fma(alpha, acc[i][j][0], beta * c.x); o.y = fma(alpha, acc[i][j][1], beta * c.y);
				} else { o.x = alpha * acc[i][j][0]; o.y = alpha * acc[i][j][1]; }
				*cp = o;
			} else {
#pragma unroll
				for (int e = 0; e < 2; e++)
					if (n + e < p.N) crow[n + e] = (beta != 0.0) ? fma(alpha, acc[i][j][e], beta * crow[n + e]) : alpha * acc[i][j][e];
			}
		}
	}
	return;
#else
#pragma unroll
	for (int i = 0; i < TM; i++) {
		const long long m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
		if (m >= p.M) continue;
		double *crow = p.C + m * p.ldc;
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const long long n = n0 + h * (BN / 2) + tx * HN;
			if (vecC && n + 1 < p.N) {
				double2 *cp = reinterpret_cast<double2 *>(crow + n);
				double2 o;
				if (beta != 0.0) {
					const double2 c = *cp;
					o.x = fma(alpha, acc[i][h * HN + 0], beta * c.x);
					o.y = fma(alpha, acc[i][h * HN + 1], beta * c.y);
				} else {
					o.x = alpha * acc[i][h * HN + 0]; o.y = alpha * acc[i][h * HN + 1];
				}
				*cp = o;
			} else {
#pragma unroll
				for (int j = 0; j < HN; j++)
					if (n + j < p.N) crow[n + j] = (beta != 0.0) ? fma(alpha, acc[i][h * HN + j], beta * crow[n + j]) : alpha * acc[i][h * HN + j];
			}
		}
	}
#endif
}

__global__ void scale_c_f64_kernel(double *C, long long ldc, int M, int N, double beta)
{
	const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	for (long long m = blockIdx.y; m < M; m += gridDim.y) {
		double *c = C + m * ldc + n;
		*c = (beta == 0.0) ? 0.0 : beta * *c;
	}
}

} // namespace

cudaError_t launch_k4_dgemm(const DProblem &p, cudaStream_t stream)
{
	const int tiles_m = (p.M + D_BM - 1) / D_BM, tiles_n = (p.N + D_BN - 1) / D_BN;
	const long long tiles = (long long)tiles_m * tiles_n;
	if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
	auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
	const bool vecA = al16(p.A) && p.lda % 2 == 0, vecB = al16(p.B) && p.ldb % 2 == 0, vecC = al16(p.C) && p.ldc % 2 == 0;
	dim3 grid((unsigned)tiles), block(D_THREADS);
#define K4_LAUNCH(AK, BKM) k4_dgemm_kernel<AK, BKM><<<grid, block, 0, stream>>>(p, tiles_m, tiles_n, vecA, vecB, vecC)
	if (p.a_kmajor) { if (p.b_kmajor) K4_LAUNCH(true, true); else K4_LAUNCH(true, false); }
	else            { if (p.b_kmajor) K4_LAUNCH(false, true); else K4_LAUNCH(false, false); }
#undef K4_LAUNCH
	return cudaGetLastError();
}

cudaError_t launch_scale_c_f64(const DProblem &p, cudaStream_t stream)
{
	if (p.M <= 0 || p.N <= 0) return cudaSuccess;
	dim3 block(256), grid((unsigned)((p.N + 255) / 256), (unsigned)min(p.M, 65535));
	scale_c_f64_kernel<<<grid, block, 0, stream>>>(p.C, p.ldc, p.M, p.N, p.beta);
	return cudaGetLastError();
}

} // namespace ugemm
