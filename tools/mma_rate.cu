// mma_rate.cu -- micro-benchmark: cycles per tcgen05.mma.kind::tf32 instruction in the 3xTF32 issue pattern
// (small*big, big*small, big*big per k-step of 8), free-running on every SM (no TMA, no transform), for
//   operand mode  SS (A and B from shared memory)  |  TS (A from tensor memory, B from shared memory)
//   cta_group     1 (UMMA M = 128)                 |  2 (UMMA M = 256 over a CTA pair)
//   UMMA N        64 .. 256
// and, optionally, with 8 extra warps per CTA streaming ld.shared/st.shared over a disjoint shared-memory region at full
// speed (what the transform warps + TMA writes do to the port in the real kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ugemm_b200/csrc -o tools/mma_rate tools/mma_rate.cu
// Output: one JSON line per configuration.
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

using namespace ugemm::ptx;

constexpr int STAGE_BYTES = 64 * 1024, STAGES = 3;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 1024;

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8])
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
	             ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one()
{
	uint32_t p;
	asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
	return p != 0;
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }

template <int CG>
__global__ void __launch_bounds__(640, 1)
mma_rate_kernel(int ts, int N, int kblocks, int noise, int nmma, int bmn, int accswap, int epi, int window, int lean, long long *out)
{
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
	const uint32_t tmem_slot = bar_base + 64;
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
	volatile uint32_t *stop_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (bar_base + 128 - smem_u32(smem_raw)));
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;

	// fill the operand stages with small finite values
	for (uint32_t i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += blockDim.x)
		sts128(smem_base + i * 16, make_float4(1.0f + (i & 7) * 0.125f, 0.5f, 0.25f, 1.5f));
	if (threadIdx.x == 0) {
		for (int s = 0; s < 8; s++) mbar_init(bar_base + 8u * s, 1);
		*stop_ptr = 0;
		fence_mbar_init();
	}
	__syncthreads();
	if (warp == 1) { tmem_alloc<CG>(tmem_slot, 512); tmem_relinquish<CG>(); }
	fence_proxy_async_smem();
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	if (warp < 4) {   // A operand area of TMEM (columns 256..511): finite values
		uint32_t r[8];
		for (int i = 0; i < 8; i++) r[i] = __float_as_uint(1.0f + 0.0625f * i);
		for (int c = 256; c < 512; c += 8) tmem_st_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
		tmem_st_wait();
	}
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	tc_fence_after();

	if (lean && warp == 1 && cta_rank == 0) {
		// lean issue loop: the warp stays converged, one elected lane issues; descriptors are advanced by adding to their low word
		if (elect_one()) {
		const uint32_t idesc = idesc_tf32(128 * CG, N, 0, bmn);
		const uint64_t dB0 = bmn ? smem_desc(0, 256, 32, 1) : smem_desc(0, 1, 64, 2), dA0 = smem_desc(0, 1, 64, 2);
		const uint32_t hiA = (uint32_t)(dA0 >> 32), hiB = (uint32_t)(dB0 >> 32), loA0 = (uint32_t)dA0, loB0 = (uint32_t)dB0;
		const uint32_t bstep = bmn ? (1024u >> 4) : (32u >> 4);
		const long long t0 = clock64();
		int s = 0;
		for (int it = 0; it < kblocks; it++) {
			if (it >= window) {
				const int w = it - window;
				while (!(CG == 2 ? mbar_try_wait_cluster(bar_base + 8u * (w & 7), (w >> 3) & 1) : mbar_try_wait(bar_base + 8u * (w & 7), (w >> 3) & 1))) {}
			}
			const uint32_t sA = (smem_base + s * STAGE_BYTES) >> 4;
			const uint32_t lAb = loA0 + sA, lBb = loB0 + sA + (16384 >> 4), lAs = lAb + (32768 >> 4), lBs = lBb + (32768 >> 4);
			const uint32_t d_tmem = tmem_base + (accswap ? (uint32_t)(((it >> 2) & 1) * 256) : 0u);
			const uint32_t tA = tmem_base + 256 + (uint32_t)(it & 1) * 128, tAs = tA + 64;
#pragma unroll
			for (int k4 = 0; k4 < 4; k4++) {
				const uint32_t first = accswap ? (((it & 3) || k4 > 0) ? 1u : 0u) : ((it > 0 || k4 > 0) ? 1u : 0u);
				const uint64_t dBb = desc64(lBb + k4 * bstep, hiB), dBs = desc64(lBs + k4 * bstep, hiB);
				if (ts) {
					mma_tf32_ts<CG>(d_tmem, tAs + k4 * 8, dBb, idesc, first);
					mma_tf32_ts<CG>(d_tmem, tA + k4 * 8, dBs, idesc, 1u);
					mma_tf32_ts<CG>(d_tmem, tA + k4 * 8, dBb, idesc, 1u);
				} else {
					const uint64_t dAb = desc64(lAb + k4 * 2, hiA), dAs = desc64(lAs + k4 * 2, hiA);
					mma_tf32_ss<CG>(d_tmem, dAs, dBb, idesc, first);
					mma_tf32_ss<CG>(d_tmem, dAb, dBs, idesc, 1u);
					mma_tf32_ss<CG>(d_tmem, dAb, dBb, idesc, 1u);
				}
			}
			mma_commit<CG>(bar_base + 8u * (it & 7));
			s = s == STAGES - 1 ? 0 : s + 1;
		}
		for (int w = kblocks > window ? kblocks - window : 0; w < kblocks; w++)
			while (!(CG == 2 ? mbar_try_wait_cluster(bar_base + 8u * (w & 7), (w >> 3) & 1) : mbar_try_wait(bar_base + 8u * (w & 7), (w >> 3) & 1))) {}
		const long long t1 = clock64();
		if (blockIdx.x < 8) out[blockIdx.x] = t1 - t0;
		*stop_ptr = 1;
		if (CG == 2) asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 1;\n\tst.shared::cluster.b32 [ra], %1;\n\t}" ::"r"(bar_base + 128), "r"(1) : "memory");
		}
		__syncwarp();
	} else if (warp == 1 && lane == 0 && cta_rank == 0) {
		const uint32_t idesc = idesc_tf32(128 * CG, N, 0, bmn);
		const long long t0 = clock64();
		for (int it = 0; it < kblocks; it++) {
			const int s = it % STAGES;
			if (it >= window) {   // at most `window` k-blocks of MMAs in flight, like a ring that waits for its stages
				const int w = it - window;
				while (!(CG == 2 ? mbar_try_wait_cluster(bar_base + 8u * (w & 7), (w >> 3) & 1) : mbar_try_wait(bar_base + 8u * (w & 7), (w >> 3) & 1))) {}
			}
			const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + 16384, sAs = sA + 32768, sBs = sB + 32768;
			const uint32_t tA = tmem_base + 256 + (uint32_t)(it & 1) * 128, tAs = tA + 64;
#pragma unroll
			for (int k4 = 0; k4 < 4; k4++) {
				const uint64_t dAb = smem_desc(sA + k4 * 32, 1, 64, 2), dAs = smem_desc(sAs + k4 * 32, 1, 64, 2);
				const uint64_t dBb = bmn ? smem_desc(sB + k4 * 1024, 256, 32, 1) : smem_desc(sB + k4 * 32, 1, 64, 2);
				const uint64_t dBs = bmn ? smem_desc(sBs + k4 * 1024, 256, 32, 1) : smem_desc(sBs + k4 * 32, 1, 64, 2);
				const uint32_t first = accswap ? (((it & 3) || k4 > 0) ? 1u : 0u) : ((it > 0 || k4 > 0) ? 1u : 0u);
				const uint32_t d_tmem = tmem_base + (accswap ? (uint32_t)(((it >> 2) & 1) * 256) : 0u);
				if (ts) {
					mma_tf32_ts<CG>(d_tmem, tAs + k4 * 8, dBb, idesc, first);
					if (nmma > 1) mma_tf32_ts<CG>(d_tmem, tA + k4 * 8, dBs, idesc, 1u);
					if (nmma > 2) mma_tf32_ts<CG>(d_tmem, tA + k4 * 8, dBb, idesc, 1u);
				} else {
					mma_tf32_ss<CG>(d_tmem, dAs, dBb, idesc, first);
					if (nmma > 1) mma_tf32_ss<CG>(d_tmem, dAb, dBs, idesc, 1u);
					if (nmma > 2) mma_tf32_ss<CG>(d_tmem, dAb, dBb, idesc, 1u);
				}
			}
			mma_commit<CG>(bar_base + 8u * (it & 7));
		}
		for (int w = kblocks > window ? kblocks - window : 0; w < kblocks; w++)
			while (!(CG == 2 ? mbar_try_wait_cluster(bar_base + 8u * (w & 7), (w >> 3) & 1) : mbar_try_wait(bar_base + 8u * (w & 7), (w >> 3) & 1))) {}
		const long long t1 = clock64();
		if (blockIdx.x < 8) out[blockIdx.x] = t1 - t0;
		*stop_ptr = 1;
		if (CG == 2) asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 1;\n\tst.shared::cluster.b32 [ra], %1;\n\t}" ::"r"(bar_base + 128), "r"(1) : "memory");
	} else if (warp >= 4 && warp < 12 && noise) {
		// shared-memory port noise: read 16 B, write 16 B per thread per step over a disjoint region (the last 16 KiB of each
		// stage is not read by the MMAs when N <= 128 ... so use the C-staging slack instead: 1 KiB after the barriers is too
		// small; reuse stage bytes the MMA never reads: B small rows >= N/CG of stage 2)
		const uint32_t region = smem_base + 2 * STAGE_BYTES + 32768 + 16384 + 8192;    // upper half of stage 2's B small: rows 64..127
		const uint32_t t = threadIdx.x - 128;                                          // 0..255
		long long n = 0;
		while (*stop_ptr == 0) {
			float4 v[4];
#pragma unroll
			for (int i = 0; i < 4; i++) v[i] = lds128(region + ((t + 256 * i) & 511) * 16);
#pragma unroll
			for (int i = 0; i < 4; i++) { v[i].x += 1.f; if (noise > 1) sts128(region + ((t + 256 * i) & 511) * 16, v[i]); }
			n++;
		}
		if (blockIdx.x == 0 && threadIdx.x == 128) out[8] = n * 4 * 256 * 16 * (noise > 1 ? 2 : 1);   // bytes moved by the noise warps
	}
	if (warp >= 12 && epi) {
		// epilogue noise: 8 warps per CTA (both CTAs of a pair) read TMEM columns with tcgen05.ld.32x32b.x16 as fast as they can
		// (epi == 2: only 1/8 of the time, roughly the duty cycle of the real promotion drains)
		const int e = warp - 12, q = e & 3, h = e >> 2;
		float acc = 0.f; long long n = 0;
		while (*stop_ptr == 0) {
#pragma unroll
			for (int g = 0; g < 8; g++) {
				float v[16];
				tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128 + g * 16), v);
#pragma unroll
				for (int i = 0; i < 16; i++) acc += v[i];
			}
			n++;
			if (epi == 2) { const long long t = clock64(); while (clock64() - t < 7 * 600) {} }
		}
		if (acc == 123.456f) out[15] = n;
		if (blockIdx.x == 0 && threadIdx.x == 12 * 32) out[9] = n * 131072;   // bytes read from TMEM by the 8 warps
	}
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	if (warp == 1) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG>
static void run(int ts, int N, int kblocks, int noise, int nmma, int bmn, int accswap, int epi, int window, int lean, long long *dout)
{
	cudaFuncSetAttribute(mma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(148); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = SMEM_BYTES;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	cudaMemset(dout, 0, 16 * sizeof(long long));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int rep = 0; rep < 2; rep++) {
		cudaEventRecord(e0);
		cudaError_t e = cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG>, ts, N, kblocks, noise, nmma, bmn, accswap, epi, window, lean, dout);
		cudaEventRecord(e1);
		if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
	}
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long h[16]; cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost);
	const double mmas = (double)kblocks * 4 * nmma;
	const double cyc = (double)h[0] / mmas;
	const double flop_per_mma = 2.0 * 128 * CG * N * 8;
	// credited 3xTF32 rate of the whole chip if every SM ran this stream: 3 MMAs = one fp32-class product
	const double tflops = (148.0 / CG) * mmas * flop_per_mma / 3.0 / (ms * 1e-3) / 1e12;
	printf("{\"exp\": \"mma_rate\", \"mode\": \"%s\", \"cg\": %d, \"N\": %d, \"mma_per_kstep\": %d, \"noise\": %d, \"kblocks\": %d, \"cycles_per_mma\": %.1f, \"floor_cycles\": %.1f, \"ms\": %.4f, \"credited_tflops_3x\": %.1f, \"clock_mhz\": %.0f, \"noise_bytes_per_clk\": %.1f, \"b_mn_major\": %d, \"accswap\": %d, \"epi\": %d, \"window\": %d, \"lean\": %d, \"tmem_ld_bytes_per_clk\": %.1f}\n",
	       ts ? "TS" : "SS", CG, N, nmma, noise, kblocks, cyc, 128.0 * N / 256.0, ms, tflops, (double)h[0] / (ms * 1e3), h[0] ? (double)h[8] / (double)h[0] : 0.0, bmn, accswap, epi, window, lean, h[0] ? (double)h[9] / (double)h[0] : 0.0);
	fflush(stdout);
}

int main(int argc, char **argv)
{
	// usage: mma_rate kblocks ts cg N noise nmma bmn accswap epi window lean  (one configuration per process: a faulting shape does not take the others down)
	auto arg = [&](int i, int d) { return argc > i ? atoi(argv[i]) : d; };
	const int kblocks = arg(1, 4096), ts = arg(2, 0), cg = arg(3, 2), N = arg(4, 256), noise = arg(5, 0), nmma = arg(6, 3);
	const int bmn = arg(7, 0), accswap = arg(8, 0), epi = arg(9, 0), window = arg(10, 2), lean = arg(11, 0);
	long long *dout; cudaMalloc(&dout, 16 * sizeof(long long));
	if (cg == 1) run<1>(ts, N, kblocks, noise, nmma, bmn, accswap, epi, window, lean, dout); else run<2>(ts, N, kblocks, noise, nmma, bmn, accswap, epi, window, lean, dout);
	return 0;
}
