// k2_simt.cu -- K2: register-blocked FFMA SGEMM for sm_100a (plain fp32, SIMT).
//
// Role (north_star): small, skinny, oddly-strided and transposed shapes that TMA cannot take
// (lda/ldb not multiples of 4, misaligned bases), and the plain-fp32 reference mode.  Replaces the
// reference's OpenCL `gemm_rnn` kernels (sgemm_ocl1.h:14-41, sgemm_ocl2.h:17-90) AND its separate
// `transpose` kernel (sgemm_ocl2.h:95-128): op(A)/op(B) are never materialised, the transposition is
// folded into the global->shared staging.
//
// Shape of the kernel: BM x BN C tile per CTA (128x128 or 64x64), BK = 16, 256 threads, each thread an
// (TM x TN) = 8x8 or 4x4 register tile split in two halves per dimension so the shared-memory reads are
// conflict-free 128-bit (64-bit) loads.  Global loads are 128-bit when the operand allows it (base 16 B
// aligned, ld % 4 == 0, quad fully in range) and guarded scalars otherwise; the next k-tile is prefetched
// into registers while the current one is multiplied (double-buffered shared memory, one barrier per tile).
// Algorithmic cost per C element: 2K flop; roofline = FP32 FFMA peak (SMs x 128 x 2 x clock).
#include "common.cuh"

namespace ugemm {

namespace {

constexpr int K2_BK = 16;
constexpr int K2_THREADS = 256;
constexpr int K2_PAD = 4;
#ifndef UGEMM_K2_PACKED
#define UGEMM_K2_PACKED 1
#endif
constexpr bool K2_PACKED_FMA = UGEMM_K2_PACKED != 0;

__device__ __forceinline__ float4 load_quad(const float *__restrict__ line, long long c, long long cmax,
                                            bool line_ok, bool vec)
{
	float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
	if (!line_ok) return v;
	if (vec && c + 3 < cmax) return __ldg(reinterpret_cast<const float4 *>(line + c));
	if (c + 0 < cmax) v.x = __ldg(line + c + 0);
	if (c + 1 < cmax) v.y = __ldg(line + c + 1);
	if (c + 2 < cmax) v.z = __ldg(line + c + 2);
	if (c + 3 < cmax) v.w = __ldg(line + c + 3);
	return v;
}

// Operand staging.  The shared tile is always [BK][BMN + PAD] (k-major rows of the M or N extent).
//   KCONTIG: global lines run along k (row-major 'N' A, or 'T' B): line index = m (or n), column = k
//   else   : global lines run along m/n (row-major 'T' A, or 'N' B): line index = k, column = m (or n)
// line-index swizzle of k-contiguous operands in shared memory (identity for the others, and for 16-wide tiles whose
// half-tiles are narrower than the swizzle): lines of rows k >= 8 are stored XOR 8
__device__ __forceinline__ constexpr int swz(int k) { return (k & 8) ? 8 : 0; }

template <int BMN, bool KCONTIG>
struct Stager {
	static constexpr int NQ = BMN * K2_BK / 4;                              // quads in the tile
	static constexpr int QUADS = (NQ + K2_THREADS - 1) / K2_THREADS;        // per thread (narrow tiles: some threads idle)
	float4 r[QUADS];

	__device__ __forceinline__ void load(const float *__restrict__ base, long long ld, long long mn0, long long mn_max,
	                                     long long k0, long long k_max, bool vec, int tid)
	{
#pragma unroll
		for (int i = 0; i < QUADS; i++) {
			int f = tid + i * K2_THREADS;
			if (NQ % K2_THREADS != 0 && f >= NQ) continue;
			if (KCONTIG) {
				int line = f / (K2_BK / 4), kq = (f % (K2_BK / 4)) * 4;
				long long mn = mn0 + line;
				r[i] = load_quad(base + mn * ld, k0 + kq, k_max, mn < mn_max, vec);
			} else {
				int k = f / (BMN / 4), q = (f % (BMN / 4)) * 4;
				long long kk = k0 + k;
				r[i] = load_quad(base + kk * ld, mn0 + q, mn_max, kk < k_max, vec);
			}
		}
	}
	// Interior fast path: the tile lies fully inside the operand, the operand is 128-bit loadable and the k-tile is
	// full, so the loads need no guards.  One pointer per thread, set once per CTA tile, advanced per k-tile; the
	// quads of one thread are a constant number of lines apart.
	const float *fp;
	long long fstep;
	__device__ __forceinline__ void fast_init(const float *__restrict__ base, long long ld, long long mn0, int tid)
	{
		if (KCONTIG) { fp = base + (mn0 + tid / (K2_BK / 4)) * ld + (tid % (K2_BK / 4)) * 4; fstep = (K2_THREADS / (K2_BK / 4)) * ld; }
		else         { fp = base + (long long)(tid / (BMN / 4)) * ld + mn0 + (tid % (BMN / 4)) * 4; fstep = (K2_THREADS / (BMN / 4)) * ld; }
	}
	__device__ __forceinline__ void fast_load(long long ld, int tid)
	{
		fp += KCONTIG ? (long long)K2_BK : K2_BK * ld;
#pragma unroll
		for (int i = 0; i < QUADS; i++) {
			if (NQ % K2_THREADS != 0 && tid + i * K2_THREADS >= NQ) continue;
			r[i] = __ldg(reinterpret_cast<const float4 *>(fp + i * fstep));
		}
	}
	// the same for an operand that may be only 4-byte aligned (odd leading dimension, sub-view): 128-bit when it can, else
	// four unguarded 32-bit loads per quad (ANYLD instantiation of the kernel)
	__device__ __forceinline__ void fast_load_any(long long ld, int tid, bool vec)
	{
		fp += KCONTIG ? (long long)K2_BK : K2_BK * ld;
#pragma unroll
		for (int i = 0; i < QUADS; i++) {
			if (NQ % K2_THREADS != 0 && tid + i * K2_THREADS >= NQ) continue;
			const float *q = fp + i * fstep;
			if (vec) r[i] = __ldg(reinterpret_cast<const float4 *>(q));
			else r[i] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
		}
	}
	__device__ __forceinline__ void store(float (*s)[BMN + K2_PAD], int tid) const
	{
#pragma unroll
		for (int i = 0; i < QUADS; i++) {
			int f = tid + i * K2_THREADS;
			if (NQ % K2_THREADS != 0 && f >= NQ) continue;
			if (KCONTIG) {
				// transposing store: a warp writes 8 lines x 4 k-quads; rows kq and kq+8 are 8*(BMN+PAD) floats = a multiple
				// of 32 banks apart, so the k >= 8 half of the tile keeps its lines XOR 8 (see swz()) and the store is conflict-free
				int line = (f / (K2_BK / 4)) ^ (BMN >= 32 ? swz(f % (K2_BK / 4) * 4) : 0), kq = (f % (K2_BK / 4)) * 4;
				s[kq + 0][line] = r[i].x; s[kq + 1][line] = r[i].y;
				s[kq + 2][line] = r[i].z; s[kq + 3][line] = r[i].w;
			} else {
				int k = f / (BMN / 4), q = (f % (BMN / 4)) * 4;
				*reinterpret_cast<float4 *>(&s[k][q]) = r[i];
			}
		}
	}
};

// ANYLD: interior tiles run the unguarded steady-state loop even when an operand is only 4-byte aligned (its quads are
// then four 32-bit loads); a separate instantiation so that the all-128-bit loop keeps its instruction stream
template <int BM, int BN, int TM, int TN, bool AK, bool BKM, bool ANYLD = false>
__global__ void __launch_bounds__(K2_THREADS, (BM >= 128 ? 2 : 3))
k2_simt_kernel(Problem p, const int tiles_m, const int tiles_n, const bool vecA, const bool vecB, const bool vecC)
{
	// strided batch: one grid.y slice per instance
	p.A += (long long)blockIdx.y * p.strideA;
	p.B += (long long)blockIdx.y * p.strideB;
	p.C += (long long)blockIdx.y * p.strideC;
	static_assert((BM / TM) * (BN / TN) == K2_THREADS, "thread tile must cover the CTA tile");
	constexpr int HM = TM / 2, HN = TN / 2;
	constexpr bool PACKED = K2_PACKED_FMA && HN % 2 == 0;   // pairs of adjacent columns inside each half of the thread tile
	__shared__ __align__(16) float As[2][K2_BK][BM + K2_PAD];
	__shared__ __align__(16) float Bs[2][K2_BK][BN + K2_PAD];

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

	float acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; i++)
#pragma unroll
		for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

	Stager<BM, AK> sa;
	Stager<BN, BKM> sb;
	const int ktiles = (p.K + K2_BK - 1) / K2_BK;

	// interior tiles of vector-loadable operands take unguarded 128-bit loads for every full k-tile (CTA-uniform test)
	const bool interior = (ANYLD || (vecA && vecB)) && m0 + BM <= p.M && n0 + BN <= p.N;
	const int fast_tiles = interior ? p.K / K2_BK : 0;
	sa.fast_init(p.A, p.lda, m0, tid);
	sb.fast_init(p.B, p.ldb, n0, tid);
	sa.load(p.A, p.lda, m0, p.M, 0, p.K, vecA, tid);
	sb.load(p.B, p.ldb, n0, p.N, 0, p.K, vecB, tid);
	sa.store(As[0], tid);
	sb.store(Bs[0], tid);
	__syncthreads();

	// first line of this thread's fragment in each half of the tile, plain and XOR 8 (k >= 8 rows of k-contiguous operands,
	// see swz()); the XOR flips bit 3 only, so the HM / HN consecutive lines of a fragment stay consecutive and aligned
	const int a_off[2] = {ty * HM, (ty * HM) ^ 8}, b_off[2] = {tx * HN, (tx * HN) ^ 8};
	auto multiply = [&](int cur) {
#pragma unroll
		for (int kk = 0; kk < K2_BK; kk++) {
			float a[TM], b[TN];
#pragma unroll
			for (int i = 0; i < HM; i++) {
				a[i] = As[cur][kk][a_off[AK && BM >= 32 && swz(kk) ? 1 : 0] + i];
				a[HM + i] = As[cur][kk][BM / 2 + a_off[AK && BM >= 32 && swz(kk) ? 1 : 0] + i];
			}
#pragma unroll
			for (int j = 0; j < HN; j++) {
				b[j] = Bs[cur][kk][b_off[BKM && BN >= 32 && swz(kk) ? 1 : 0] + j];
				b[HN + j] = Bs[cur][kk][BN / 2 + b_off[BKM && BN >= 32 && swz(kk) ? 1 : 0] + j];
			}
			if constexpr (PACKED) {
				// Blackwell packed fp32: one FFMA2 updates two adjacent columns (a 64-bit register pair), which halves the
				// issue slots per flop and reads every operand as an even/odd register pair (no bank-conflict lottery)
#pragma unroll
				for (int i = 0; i < TM; i++) {
					const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
					for (int j = 0; j < TN; j += 2) {
						const float2 r = __ffma2_rn(aa, make_float2(b[j], b[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
						acc[i][j] = r.x; acc[i][j + 1] = r.y;
					}
				}
			} else {
#pragma unroll
				for (int i = 0; i < TM; i++)
#pragma unroll
					for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
			}
		}
	};

	int t = 0;
	// steady state of interior tiles: nothing but 128-bit loads, the FFMA block, the shared-memory stores and one barrier
	for (; t + 1 < fast_tiles; t++) {
		const int cur = t & 1;
		if (ANYLD) { sa.fast_load_any(p.lda, tid, vecA); sb.fast_load_any(p.ldb, tid, vecB); }
		else { sa.fast_load(p.lda, tid); sb.fast_load(p.ldb, tid); }
		multiply(cur);
		sa.store(As[cur ^ 1], tid);
		sb.store(Bs[cur ^ 1], tid);
		__syncthreads();
	}
	// edge tiles, scalar-only operands and the K tail: guarded loads
	for (; t < ktiles; t++) {
		const int cur = t & 1;
		if (t + 1 < ktiles) {
			sa.load(p.A, p.lda, m0, p.M, (long long)(t + 1) * K2_BK, p.K, vecA, tid);
			sb.load(p.B, p.ldb, n0, p.N, (long long)(t + 1) * K2_BK, p.K, vecB, tid);
		}
		multiply(cur);
		if (t + 1 < ktiles) {
			sa.store(As[cur ^ 1], tid);
			sb.store(Bs[cur ^ 1], tid);
		}
		__syncthreads();
	}

	// fused epilogue: C = alpha*acc + beta*C (C never read when beta == 0), ld padding never touched
	const float alpha = p.alpha, beta = p.beta, slope = p.slope;
	const bool post = p.bias != nullptr || slope != 1.f;   // bias[m] + LeakyReLU of the convolution callers
#pragma unroll
	for (int i = 0; i < TM; i++) {
		const long long m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
		if (m >= p.M) continue;
		float *crow = p.C + m * p.ldc;
		const float bm = p.bias ? __ldg(p.bias + m) : 0.f;
		auto act = [&](float x) { x += bm; return x > 0.f ? x : x * slope; };
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const long long n = n0 + h * (BN / 2) + tx * HN;
			if (HN == 4 && vecC && n + 3 < p.N) {
				float4 o;
				float4 *cp = reinterpret_cast<float4 *>(crow + n);
				if (beta != 0.f) {
					float4 c = *cp;
					o.x = fmaf(alpha, acc[i][h * HN + 0], beta * c.x);
					o.y = fmaf(alpha, acc[i][h * HN + 1], beta * c.y);
					o.z = fmaf(alpha, acc[i][h * HN + 2], beta * c.z);
					o.w = fmaf(alpha, acc[i][h * HN + 3], beta * c.w);
				} else {
					o.x = alpha * acc[i][h * HN + 0]; o.y = alpha * acc[i][h * HN + 1];
					o.z = alpha * acc[i][h * HN + 2]; o.w = alpha * acc[i][h * HN + 3];
				}
				if (post) { o.x = act(o.x); o.y = act(o.y); o.z = act(o.z); o.w = act(o.w); }
				*cp = o;
			} else {
#pragma unroll
				for (int j = 0; j < HN; j++) {
					if (n + j < p.N) {
						float v = alpha * acc[i][h * HN + j];
						if (beta != 0.f) v = fmaf(alpha, acc[i][h * HN + j], beta * crow[n + j]);
						crow[n + j] = post ? act(v) : v;
					}
				}
			}
		}
	}
}

__global__ void scale_c_kernel(float *C, long long ldc, int M, int N, float beta)
{
	long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long m = blockIdx.y;
	if (n >= N) return;
	for (; m < M; m += gridDim.y) {
		float *c = C + m * ldc + n;
		*c = (beta == 0.f) ? 0.f : beta * *c;
	}
}

// col[(c*k*k + ki*k + kj) * (Ho*Wo) + (io*Wo + jo)] = im[c][io*stride - pad + ki][jo*stride - pad + kj] (0 outside the image):
// the layout of the reference's im2col (sgemm_ocl1.h:81-119, sgemm_gl1.h:166-190).  One thread per column-matrix
// element, pixels fastest, so both the writes and (for stride 1) the reads are coalesced; HBM-bound by the
// 4*C*k*k*Ho*Wo bytes it writes.
__global__ void im2col_kernel(const float *__restrict__ im, int channels, int height, int width, int k, int pad, int stride,
                              int ho, int wo, float *__restrict__ col)
{
	const long long npix = (long long)ho * wo, total = (long long)channels * k * k * npix;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
		const long long row = idx / npix, pix = idx - row * npix;
		const int kj = (int)(row % k), ki = (int)((row / k) % k), c = (int)(row / ((long long)k * k));
		const int io = (int)(pix / wo), jo = (int)(pix - (long long)io * wo);
		const int i = io * stride - pad + ki, j = jo * stride - pad + kj;
		col[idx] = (i >= 0 && j >= 0 && i < height && j < width) ? __ldg(im + ((long long)c * height + i) * width + j) : 0.f;
	}
}

template <int BM, int BN, int TM, int TN>
cudaError_t launch_cfg(const Problem &p, cudaStream_t stream)
{
	const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
	const long long tiles = (long long)tiles_m * tiles_n;
	if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
	auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
	const bool multi = p.batch > 1;
	const bool vecA = al16(p.A) && (p.lda % 4 == 0) && (!multi || p.strideA % 4 == 0);
	const bool vecB = al16(p.B) && (p.ldb % 4 == 0) && (!multi || p.strideB % 4 == 0);
	const bool vecC = al16(p.C) && (p.ldc % 4 == 0) && (!multi || p.strideC % 4 == 0);
	if (p.batch > 65535) return cudaErrorInvalidConfiguration;
	dim3 grid((unsigned)tiles, (unsigned)(multi ? p.batch : 1)), block(K2_THREADS);
#define K2_LAUNCH(AK, BKM) do { \
	if (BM == 128 && BN == 128 && !(vecA && vecB)) \
		k2_simt_kernel<BM, BN, TM, TN, AK, BKM, (BM == 128 && BN == 128)><<<grid, block, 0, stream>>>(p, tiles_m, tiles_n, vecA, vecB, vecC); \
	else k2_simt_kernel<BM, BN, TM, TN, AK, BKM><<<grid, block, 0, stream>>>(p, tiles_m, tiles_n, vecA, vecB, vecC); } while (0)
	if (p.a_kmajor) { if (p.b_kmajor) K2_LAUNCH(true, true); else K2_LAUNCH(true, false); }
	else            { if (p.b_kmajor) K2_LAUNCH(false, true); else K2_LAUNCH(false, false); }
#undef K2_LAUNCH
	return cudaGetLastError();
}

} // namespace

cudaError_t launch_k2_simt(const Problem &p, cudaStream_t stream, int sm_count)
{
	// 128x128 tiles once they fill the machine (>= one CTA per SM); 64x64 below that so that small and
	// skinny problems still spread over the 148 SMs.
	// narrow N (the tall-skinny, HBM-bound corner): 256-row tiles that are only as wide as the problem, so the FMA work
	// wasted on columns >= N does not turn a memory-bound shape into a compute-bound one
	if (p.N <= 16 && p.M >= 256) return launch_cfg<256, 16, 8, 2>(p, stream);
	if (p.N <= 32 && p.M >= 256) return launch_cfg<256, 32, 8, 4>(p, stream);
	if (p.N <= 64 && (long long)(p.M / 256) * (p.batch > 0 ? p.batch : 1) >= sm_count) return launch_cfg<256, 64, 8, 8>(p, stream);
	const long long big_tiles = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128) * (p.batch > 0 ? p.batch : 1);
	if (big_tiles >= sm_count) return launch_cfg<128, 128, 8, 8>(p, stream);
	return launch_cfg<64, 64, 4, 4>(p, stream);
}

cudaError_t launch_im2col(const float *im, int channels, int height, int width, int k, int pad, int stride, float *col, cudaStream_t stream)
{
	const int ho = (height + 2 * pad - k) / stride + 1, wo = (width + 2 * pad - k) / stride + 1;
	const long long total = (long long)channels * k * k * ho * wo;
	if (total <= 0) return cudaSuccess;
	long long blocks = (total + 255) / 256;
	if (blocks > 148LL * 64) blocks = 148LL * 64;
	im2col_kernel<<<(unsigned)blocks, 256, 0, stream>>>(im, channels, height, width, k, pad, stride, ho, wo, col);
	return cudaGetLastError();
}

cudaError_t launch_scale_c(const Problem &p, cudaStream_t stream)
{
	if (p.M <= 0 || p.N <= 0) return cudaSuccess;
	dim3 block(256), grid((unsigned)((p.N + 255) / 256), (unsigned)min(p.M, 65535));
	for (int b = 0; b < (p.batch > 0 ? p.batch : 1); b++) {
		scale_c_kernel<<<grid, block, 0, stream>>>(p.C + (long long)b * p.strideC, p.ldc, p.M, p.N, p.beta);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

} // namespace ugemm
