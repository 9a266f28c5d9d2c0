// k3_level12.cu -- the HBM-bound level-1/level-2 companions of the SGEMM path: SAXPY and SGEMV for sm_100a.
//
// Replaces the reference's CPU loops saxpy_cpu (ugemm.h:75-86) / saxpy_avx (ugemm.h:58-73), sgemv_cpu
// (ugemm.h:124-150) and the OpenCL Xaxpy kernel of saxpy_ocl.c:129-157.  Neither has any reuse, so both are bound by
// HBM bandwidth and are written for exactly that: 128-bit coalesced accesses, several independent loads in flight per
// thread, grids sized in multiples of the SM count, no shared-memory staging of the matrix (every element is used once).
//   saxpy   algorithmic bytes 12*N (read x, read y, write y), 2*N flop
//   sgemv   algorithmic bytes 4*M*N (+ vectors),              2*M*N flop
#include "common.cuh"

namespace ugemm {

namespace {

constexpr int L12_THREADS = 256;

// y[i*incy] += alpha * x[i*incx]; one fused multiply-add per element like the contracted reference loop
__global__ void __launch_bounds__(L12_THREADS)
saxpy_vec_kernel(long long n4, float alpha, const float4 *__restrict__ x, float4 *__restrict__ y)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	// two independent quads per trip: 64 B of loads in flight per thread before the first store
	for (; i + stride < n4; i += 2 * stride) {
		const float4 x0 = __ldg(x + i), x1 = __ldg(x + i + stride);
		float4 y0 = y[i], y1 = y[i + stride];
		y0.x = fmaf(alpha, x0.x, y0.x); y0.y = fmaf(alpha, x0.y, y0.y); y0.z = fmaf(alpha, x0.z, y0.z); y0.w = fmaf(alpha, x0.w, y0.w);
		y1.x = fmaf(alpha, x1.x, y1.x); y1.y = fmaf(alpha, x1.y, y1.y); y1.z = fmaf(alpha, x1.z, y1.z); y1.w = fmaf(alpha, x1.w, y1.w);
		y[i] = y0; y[i + stride] = y1;
	}
	if (i < n4) {
		const float4 x0 = __ldg(x + i);
		float4 y0 = y[i];
		y0.x = fmaf(alpha, x0.x, y0.x); y0.y = fmaf(alpha, x0.y, y0.y); y0.z = fmaf(alpha, x0.z, y0.z); y0.w = fmaf(alpha, x0.w, y0.w);
		y[i] = y0;
	}
}

__global__ void __launch_bounds__(L12_THREADS)
saxpy_strided_kernel(long long first, long long n, float alpha, const float *__restrict__ x, long long incx, float *__restrict__ y, long long incy)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		y[i * incy] = fmaf(alpha, __ldg(x + i * incx), y[i * incy]);
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Rows of A contiguous along the summed index (the reference's trans != 'N' case, A[n + m*lda]): TPR threads share one
// row, each walking it in 128-bit steps with 4 independent partial sums, then a shuffle (+ shared-memory) reduction.
// y[m*incy] = alpha*sum + beta*y[m*incy]; y is not read when beta == 0.
template <int TPR>
__global__ void __launch_bounds__(L12_THREADS)
sgemv_rows_kernel(int M, int N, float alpha, const float *__restrict__ A, long long lda, const float *__restrict__ x, long long incx,
                  float beta, float *__restrict__ y, long long incy, bool vec)
{
	constexpr int RPB = L12_THREADS / TPR;                // rows per block
	__shared__ float part[L12_THREADS / 32];
	const int t = threadIdx.x % TPR;
	const long long m = (long long)blockIdx.x * RPB + threadIdx.x / TPR;
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
	if (m < M) {
		const float *row = A + m * lda;
		if (vec) {
			const float4 *row4 = reinterpret_cast<const float4 *>(row);
			const float4 *x4 = reinterpret_cast<const float4 *>(x);
			const int n4 = N / 4;
			int q = t;
			for (; q + TPR < n4; q += 2 * TPR) {
				const float4 a0 = __ldg(row4 + q), a1 = __ldg(row4 + q + TPR);
				const float4 b0 = __ldg(x4 + q), b1 = __ldg(x4 + q + TPR);
				s0 = fmaf(a0.x, b0.x, s0); s1 = fmaf(a0.y, b0.y, s1); s2 = fmaf(a0.z, b0.z, s2); s3 = fmaf(a0.w, b0.w, s3);
				s0 = fmaf(a1.x, b1.x, s0); s1 = fmaf(a1.y, b1.y, s1); s2 = fmaf(a1.z, b1.z, s2); s3 = fmaf(a1.w, b1.w, s3);
			}
			if (q < n4) {
				const float4 a0 = __ldg(row4 + q), b0 = __ldg(x4 + q);
				s0 = fmaf(a0.x, b0.x, s0); s1 = fmaf(a0.y, b0.y, s1); s2 = fmaf(a0.z, b0.z, s2); s3 = fmaf(a0.w, b0.w, s3);
			}
			for (int n = n4 * 4 + t; n < N; n += TPR) s0 = fmaf(__ldg(row + n), __ldg(x + n), s0);
		} else {
			for (int n = t; n < N; n += TPR) s0 = fmaf(__ldg(row + n), __ldg(x + (long long)n * incx), s0);
		}
	}
	float s = warp_sum((s0 + s1) + (s2 + s3));
	if (TPR > 32) {
		if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
		__syncthreads();
		if (threadIdx.x == 0) {
			s = 0.f;
#pragma unroll
			for (int w = 0; w < L12_THREADS / 32; w++) s += part[w];
		}
	}
	if (t == 0 && m < M) {
		float *yp = y + m * incy;
		*yp = (beta == 0.f) ? alpha * s : fmaf(alpha, s, beta * *yp);
	}
}

// Columns of A contiguous along the OUTPUT index (the reference's trans == 'N' case, A[m + n*lda]): a lane owns V
// consecutive m (V = 4: one 128-bit load per n when the layout allows, else V = 1), a warp 32*V of them (one or four
// coalesced 128-byte lines per n); the 8 warps of a block take n = w, w+8, ... of the block's n range and meet in
// shared memory.  Four n per trip keep four lines in flight per lane.  When M alone cannot fill the machine the n range
// is cut over grid.y and the slices' partial sums go to a scratch array [slices][M] that sgemv_finish_kernel adds up in
// a fixed order (no atomics: the result does not depend on scheduling).
template <int V>
__global__ void __launch_bounds__(L12_THREADS)
sgemv_cols_kernel(int M, int N, int n_per_slice, float alpha, const float *__restrict__ A, long long lda, const float *__restrict__ x,
                  long long incx, float beta, float *__restrict__ y, long long incy, float *__restrict__ partial)
{
	constexpr int WARPS = L12_THREADS / 32;
	__shared__ float part[WARPS][32 * V];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const long long m = ((long long)blockIdx.x * 32 + lane) * V;
	const int n0 = blockIdx.y * n_per_slice, n1 = min(N, n0 + n_per_slice);
	float s[4][V];
#pragma unroll
	for (int u = 0; u < 4; u++)
#pragma unroll
		for (int v = 0; v < V; v++) s[u][v] = 0.f;
	if (m < M) {
		const float *col = A + m;
		auto ld = [&](int n, float (&a)[V]) {
			if (V == 4) {
				const float4 q = __ldg(reinterpret_cast<const float4 *>(col + (long long)n * lda));
				a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
			} else {
				a[0] = __ldg(col + (long long)n * lda);
			}
		};
		int n = n0 + w;
		for (; n + 3 * WARPS < n1; n += 4 * WARPS) {
			float a[4][V], xv[4];
#pragma unroll
			for (int u = 0; u < 4; u++) { ld(n + u * WARPS, a[u]); xv[u] = __ldg(x + (long long)(n + u * WARPS) * incx); }
#pragma unroll
			for (int u = 0; u < 4; u++)
#pragma unroll
				for (int v = 0; v < V; v++) s[u][v] = fmaf(a[u][v], xv[u], s[u][v]);
		}
		for (; n < n1; n += WARPS) {
			float a[V];
			ld(n, a);
			const float xv = __ldg(x + (long long)n * incx);
#pragma unroll
			for (int v = 0; v < V; v++) s[0][v] = fmaf(a[v], xv, s[0][v]);
		}
	}
#pragma unroll
	for (int v = 0; v < V; v++) part[w][lane * V + v] = (s[0][v] + s[1][v]) + (s[2][v] + s[3][v]);
	__syncthreads();
	// one thread per output of the block adds the 8 warps' sums in warp order
	for (int o = threadIdx.x; o < 32 * V; o += L12_THREADS) {
		const long long mo = (long long)blockIdx.x * 32 * V + o;
		if (mo >= M) continue;
		float t = 0.f;
#pragma unroll
		for (int i = 0; i < WARPS; i++) t += part[i][o];
		if (partial) partial[(long long)blockIdx.y * M + mo] = t;
		else {
			float *yp = y + mo * incy;
			*yp = (beta == 0.f) ? alpha * t : fmaf(alpha, t, beta * *yp);
		}
	}
}

__global__ void __launch_bounds__(L12_THREADS)
sgemv_finish_kernel(int M, int slices, float alpha, const float *__restrict__ partial, float beta, float *__restrict__ y, long long incy)
{
	const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= M) return;
	float t = 0.f;
	for (int sl = 0; sl < slices; sl++) t += partial[(long long)sl * M + m];
	float *yp = y + m * incy;
	*yp = (beta == 0.f) ? alpha * t : fmaf(alpha, t, beta * *yp);
}

bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

} // namespace

cudaError_t launch_saxpy(long long n, float alpha, const float *x, long long incx, float *y, long long incy, cudaStream_t stream, int sm_count)
{
	if (n <= 0) return cudaSuccess;
	long long done = 0;
	if (incx == 1 && incy == 1 && al16(x) && al16(y) && n >= 4) {
		const long long n4 = n / 4;
		long long blocks = (n4 + 2 * L12_THREADS - 1) / (2 * L12_THREADS);
		const long long cap = (long long)sm_count * 8;           // 8 resident CTAs of 256 threads per SM
		if (blocks > cap) blocks = cap;
		saxpy_vec_kernel<<<(unsigned)blocks, L12_THREADS, 0, stream>>>(n4, alpha, reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(y));
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
		done = n4 * 4;
	}
	if (done < n) {
		long long blocks = (n - done + L12_THREADS - 1) / L12_THREADS;
		const long long cap = (long long)sm_count * 8;
		if (blocks > cap) blocks = cap;
		saxpy_strided_kernel<<<(unsigned)blocks, L12_THREADS, 0, stream>>>(done, n, alpha, x, incx, y, incy);
	}
	return cudaGetLastError();
}

// rows_contiguous: A[n + m*lda] (trans != 'N' in the reference); otherwise A[m + n*lda]
cudaError_t launch_sgemv(bool rows_contiguous, int M, int N, float alpha, const float *A, long long lda, const float *x, long long incx,
                         float beta, float *y, long long incy, cudaStream_t stream, int sm_count)
{
	if (M <= 0) return cudaSuccess;
	if (rows_contiguous) {
		const bool vec = incx == 1 && al16(A) && al16(x) && lda % 4 == 0;
		// a warp per row once that fills the machine (8 rows per block); a whole block per row for few, long rows
		if ((long long)M >= (long long)sm_count * 8 * 4 || N < 2048)
			sgemv_rows_kernel<32><<<(unsigned)((M + 7) / 8), L12_THREADS, 0, stream>>>(M, N, alpha, A, lda, x, incx, beta, y, incy, vec);
		else
			sgemv_rows_kernel<L12_THREADS><<<(unsigned)M, L12_THREADS, 0, stream>>>(M, N, alpha, A, lda, x, incx, beta, y, incy, vec);
	} else {
		const bool vec = al16(A) && lda % 4 == 0 && M % 4 == 0;
		const int per_block = vec ? 128 : 32;
		const long long bx = ((long long)M + per_block - 1) / per_block;
		// enough blocks for ~4 per SM; a slice keeps >= 256 n so the slicing overhead stays small
		long long slices = ((long long)sm_count * 4 + bx - 1) / bx;
		if (slices > (N + 255) / 256) slices = (N + 255) / 256;
		if (slices < 1) slices = 1;
		if (slices > 65535) slices = 65535;
		const int n_per_slice = (int)(((long long)N + slices - 1) / slices);
		slices = n_per_slice > 0 ? ((long long)N + n_per_slice - 1) / n_per_slice : 1;
		if (slices < 1) slices = 1;
		float *partial = nullptr;
		if (slices > 1 && cudaMallocAsync(reinterpret_cast<void **>(&partial), (size_t)slices * M * sizeof(float), stream) != cudaSuccess) {
			cudaGetLastError();
			slices = 1;      // no scratch: one slice per block column, still correct
		}
		const int nps = slices > 1 ? n_per_slice : (N > 0 ? N : 1);
		dim3 grid((unsigned)bx, (unsigned)slices);
		if (vec) sgemv_cols_kernel<4><<<grid, L12_THREADS, 0, stream>>>(M, N, nps, alpha, A, lda, x, incx, beta, y, incy, slices > 1 ? partial : nullptr);
		else     sgemv_cols_kernel<1><<<grid, L12_THREADS, 0, stream>>>(M, N, nps, alpha, A, lda, x, incx, beta, y, incy, slices > 1 ? partial : nullptr);
		cudaError_t e = cudaGetLastError();
		if (slices > 1) {
			if (e == cudaSuccess) {
				sgemv_finish_kernel<<<(unsigned)((M + L12_THREADS - 1) / L12_THREADS), L12_THREADS, 0, stream>>>(M, (int)slices, alpha, partial, beta, y, incy);
				e = cudaGetLastError();
			}
			cudaFreeAsync(partial, stream);
		}
		return e;
	}
	return cudaGetLastError();
}

} // namespace ugemm
