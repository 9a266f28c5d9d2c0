// k1_tcgen05.cu -- K1: 3xTF32 error-compensated SGEMM on the 5th-gen tensor cores (tcgen05 / TMEM / TMA): host side, stream-K fix-up
// pass, hardware probe.  The kernels themselves are in the headers this file includes (one translation unit):
//     k1_common.cuh   stage geometry, parameter block, work items, tile-index ring, 3xTF32 split, epilogue pieces
//     k1_ts.cuh       k1ts_kernel: the production kernel, op(A) in tensor memory, 64-column accumulator slices
//     k1_ss.cuh       k1_3xtf32_kernel: the round-1 kernel, both operands from shared memory (A/B runs, RNA-split experiment)
//
// C = alpha * op(A) op(B) + beta * C with fp32-class accuracy from TF32 tensor-core products:
//     a = a_big + a_small,  a_big = tf32(a) (top 19 bits),  a_small = a - a_big  (exact in fp32)
//     a*b ~= a_small*b_big + a_big*b_small + a_big*b_big        (a_small*b_small ~ 2^-22 dropped)
// Three tcgen05.mma per k-step, credited as 2*M*N*K flop (effective roofline = dense TF32 peak / 3).
//
// Replaces the reference's blocked inner loops (sgemm_avx256.h:20-390, gemm_cpu.h:96-282, the OpenCL
// gemm_fast / gemm_rnn kernels sgemm_ocl.h:444-538, sgemm_ocl2.h:17-90) for TMA-eligible problems.
//
// Common structure of both kernels (one CTA per SM, optionally paired as a 2-CTA cluster issuing cta_group::2 MMAs; 640 threads):
//   warp 0      TMA producer: raw fp32 tiles of op(A) (128 rows) and op(B) (128 rows) per k-block of 32, 128B-swizzled, into a
//               shared-memory ring.  K-major operands: {32 k x rows} boxes, SWIZZLE_128B.  MN-major operands (transA=='T' /
//               transB=='N'): {32 mn x 32 k} boxes, SWIZZLE_128B_ATOM_32B (the only MN-major layout tcgen05 accepts for 32-bit
//               types).  Ragged M/N/K edges are zero-filled by TMA out-of-bounds handling.  CONV instantiation: the B boxes are
//               gathered from a channels-last image with 4-D coordinates (implicit im2col).
//   warp 1      MMA issuer (leader CTA only): per k-step of 8: small*big, big*small, big*big into TMEM accumulators;
//               tcgen05.commit releases the stage.
//   warp 2      work scheduler (leader CTA only): claims tile indices from a global atomic counter and publishes them to every role
//               of both CTAs through a 4-deep shared-memory ring; stream-K tail ranges for the last partial round.
//   warps 4-11  transform: the "small" parts (TS: op(A) raw + small straight into TMEM with tcgen05.st, B small into shared memory;
//               SS: both small copies into shared memory), fence.proxy.async, signal.
//   warps 12-19 epilogue: tcgen05.ld the accumulator, promote partial sums every kc_blocks k-blocks into fp32 registers with
//               round-to-nearest adds, then fused alpha/beta(/bias/LeakyReLU) and TMA box stores (row-strided stores for a C that
//               TMA cannot address) -- or, for a stream-K part, raw partial sums into the workspace that k1_tail_fixup_kernel adds up.
//   Persistent: tiles are handed out in an L2-friendly grouped order (8 m-tiles share an n sweep).
#include "k1_common.cuh"
#include "k1_ss.cuh"
#include "k1_ts.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ugemm {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Stream-K fix-up: C tile = alpha * (sum of the tile's partial-sum parts, in range order) + beta * C (+ bias, LeakyReLU).
// FIXUP_SPLIT CTAs per tail tile; the parts are L2-resident (just written by K1).  Deterministic: no atomics, fixed order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FIXUP_ROWS = 4;        // tile rows per CTA: every thread owns ONE 16-byte quad of C, so that all loads of the pass are in flight at once
template <int CG>
__global__ void __launch_bounds__(256)
k1_tail_fixup_kernel(const K1Params P)
{
	constexpr int TM_ = 128 * CG, TN_ = 128 * CG, Q = TN_ / 4;
	static_assert(FIXUP_ROWS * Q % 256 == 0 || FIXUP_ROWS * Q <= 256, "one quad per thread");
	griddep_wait();        // launched with programmatic stream serialization: the parts are complete once the GEMM kernel has finished
	const int r = blockIdx.x;
	int tm, tn;
	const int inst = (P.sk_full + r) / P.tiles_per_batch;
	decode_tile(P.sk_full + r - inst * P.tiles_per_batch, P.tiles_m, P.tiles_n, tm, tn, P.group);
	const bool conv = P.cv_wp > 0;
	const int r_lo = (r * P.sk_nch) / P.sk_q, r_hi = ((r + 1) * P.sk_nch - 1) / P.sk_q;   // chunk ranges that hold a part of this tile
	const float alpha = P.alpha, beta = P.beta, slope = P.slope;
	const bool post = P.bias != nullptr || slope != 1.f;
	// a range's part of this tile is its first item (h = 0) if the range starts inside the tile, else its second
	auto slot_of = [&](int rr) { return 2 * rr + (((rr * P.sk_q) / P.sk_nch == r) ? 0 : 1); };
	for (int idx = threadIdx.x; idx < FIXUP_ROWS * Q; idx += blockDim.x) {
		const int row = blockIdx.y * FIXUP_ROWS + idx / Q, c4 = (idx % Q) * 4;
		const long long gm = (long long)tm * TM_ + row, gn = (long long)tn * TN_ + c4;
		if (gm >= P.M || gn >= P.N) continue;
		const float *part = P.sk_ws + (long long)row * TN_ + c4;
		// the parts are added in range order (bit-reproducible); four independent loads at a time
		float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
		for (int rr = r_lo; rr <= r_hi; rr += 4) {
			float4 v[4];
#pragma unroll
			for (int u = 0; u < 4; u++)
				v[u] = rr + u <= r_hi ? *reinterpret_cast<const float4 *>(part + (long long)slot_of(rr + u) * (TM_ * TN_)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
			for (int u = 0; u < 4; u++)
				if (rr + u <= r_hi) { sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w; }
		}
		float *cp = P.C + (long long)inst * P.strideC + gm * P.ldc + gn;
		int valid = 4;                              // elements of this quad that exist in C
		bool vec = P.vecC;
		if (conv) {                                 // padded column index -> (output row, output column); a quad never straddles rows
			const int io = (int)(gn / P.cv_wp), jo = (int)(gn - (long long)io * P.cv_wp);
			valid = io < P.cv_ho ? P.cv_wo - jo : 0;
			const int off = io * P.cv_wo + jo;
			cp = P.C + (long long)inst * P.strideC + gm * (long long)P.cv_npix + off;
			vec = vec && (off & 3) == 0;
		} else if (gn + 3 >= P.N) valid = (int)(P.N - gn);
		if (valid <= 0) continue;
		const float bm = P.bias ? __ldg(P.bias + gm) : 0.f;
		auto fin = [&](float acc, float cold) {
			float o = beta != 0.f ? fmaf(alpha, acc, beta * cold) : alpha * acc;
			if (post) { o += bm; o = o > 0.f ? o : o * slope; }
			return o;
		};
		if (vec && valid >= 4) {
			float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
			if (beta != 0.f) c = *reinterpret_cast<const float4 *>(cp);
			*reinterpret_cast<float4 *>(cp) = make_float4(fin(sum.x, c.x), fin(sum.y, c.y), fin(sum.z, c.z), fin(sum.w, c.w));
		} else {
			const float sv[4] = {sum.x, sum.y, sum.z, sum.w};
			for (int e = 0; e < 4; e++)
				if (e < valid) cp[e] = fin(sv[e], beta != 0.f ? cp[e] : 0.f);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Probe: one CTA, manual K-major SWIZZLE_128B staging (no TMA), `ksteps` chained 128x16x8 TF32 MMAs.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_tf32_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int ksteps, unsigned *diag)
{
	__shared__ __align__(1024) uint8_t sA[128 * 128];
	__shared__ __align__(1024) uint8_t sB[16 * 128];
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t tmem_slot;
	const int tid = threadIdx.x, warp = tid >> 5;
	const int K = 8 * ksteps; // <= 32
	// K-major SW128: element (r,k) at r*128 + ((k/4) ^ (r%8))*16 + (k%4)*4
	for (int idx = tid; idx < 128 * 32; idx += 128) {
		int r = idx / 32, k = idx % 32;
		float v = k < K ? A[r * K + k] : 0.f;
		*reinterpret_cast<float *>(sA + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) = v;
	}
	for (int idx = tid; idx < 16 * 32; idx += 128) {
		int r = idx / 32, k = idx % 32;
		float v = k < K ? B[r * K + k] : 0.f;
		*reinterpret_cast<float *>(sB + r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4) = v;
	}
	if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
	if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_slot), 32); tmem_relinquish<1>(); }
	fence_proxy_async_smem();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = tmem_slot;
	if (tid == 0) {
		const uint32_t idesc = idesc_tf32(128, 16, 0, 0);
		for (int k = 0; k < ksteps; k++) {
			const uint64_t da = smem_desc(smem_u32(sA) + k * 32, 1, 64, 2);
			const uint64_t db = smem_desc(smem_u32(sB) + k * 32, 1, 64, 2);
			mma_tf32_ss<1>(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
		}
		mma_commit<1>(smem_u32(&bar));
	}
	mbar_wait(smem_u32(&bar), 0, diag, 9);
	tc_fence_after();
	float v[16];
	tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
	for (int i = 0; i < 16; i++) D[tid * 16 + i] = v[i];
	tc_fence_before();
	__syncthreads();
	if (warp == 0) tmem_dealloc<1>(tmem_base, 32);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
	static EncodeTiledFn fn = nullptr;
	static std::once_flag once;
	std::call_once(once, [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
		    q == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(p);
	});
	return fn;
}

// Tensor-map cache.  A descriptor is a pure function of (base, extents, pitches, box, swizzle); GEMM callers launch the same
// operands again and again (every step of a benchmark loop, every K slab of the sharded driver, every instance walk of the
// harness), so the encoded 128-byte maps are kept in a small per-thread table (no lock, no sharing) keyed by exactly those
// arguments and cuTensorMapEncodeTiled runs only on a miss.
struct MapKey {
	const void *base; unsigned long long d0, d1, d2, d3, s0, s1, s2; unsigned b0, b1, b2, b3, e1, rank, swizzle, l2;
	bool operator==(const MapKey &o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
struct MapCache {
	static constexpr int N = 64;
	MapKey key[N]; CUtensorMap map[N]; bool used[N]; unsigned long long hits, misses;
	MapCache() : hits(0), misses(0) { memset(key, 0, sizeof key); memset(used, 0, sizeof used); }
};
bool cached_encode(CUtensorMap *out, unsigned rank, const void *base, const cuuint64_t *gdim, const cuuint64_t *gstride, const cuuint32_t *box,
                   const cuuint32_t *estr, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2)
{
	EncodeTiledFn fn = encode_fn();
	if (!fn) return false;
	thread_local MapCache cache;
	MapKey k;
	memset(&k, 0, sizeof k);            // padding bytes too: the key is compared with memcmp
	k.base = base; k.rank = rank; k.swizzle = (unsigned)sw; k.l2 = (unsigned)l2;
	k.d0 = gdim[0]; k.d1 = gdim[1]; k.d2 = rank > 2 ? gdim[2] : 0; k.d3 = rank > 3 ? gdim[3] : 0;
	k.s0 = gstride[0]; k.s1 = rank > 2 ? gstride[1] : 0; k.s2 = rank > 3 ? gstride[2] : 0;
	k.b0 = box[0]; k.b1 = box[1]; k.b2 = rank > 2 ? box[2] : 0; k.b3 = rank > 3 ? box[3] : 0; k.e1 = estr[1];
	unsigned long long h = reinterpret_cast<uintptr_t>(base) >> 4;
	h = mix64(h ^ (k.d0 * 0x9E3779B97F4A7C15ull) ^ (k.d1 << 17) ^ (k.s0 << 29) ^ (k.s1 << 3) ^ ((unsigned long long)k.b1 << 50) ^ k.swizzle);
	const int slot = (int)(h % MapCache::N);
	if (cache.used[slot] && cache.key[slot] == k) { *out = cache.map[slot]; cache.hits++; return true; }
	if (fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
	       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
		return false;
	cache.key[slot] = k; cache.map[slot] = *out; cache.used[slot] = true; cache.misses++;
	return true;
}

// K-major operand: `rows` lines of `K` contiguous fp32, pitch ld  -> dims {K, rows, batch}, box {32, 128, 1}, SWIZZLE_128B
// MN-major operand: `K` lines of `rows` contiguous fp32, pitch ld -> dims {rows, K, batch}, box {32, 32, 1}, SWIZZLE_128B_ATOM_32B
// The third dimension walks the strided batch (extent 1 for a plain GEMM).
bool make_operand_map(CUtensorMap *map, const float *base, long long rows, long long K, long long ld, bool kmajor,
                      int batch, long long stride, int box_rows = ROWS)
{
	cuuint64_t gdim[3], gstride[2];
	cuuint32_t box[3], estr[3] = {1, 1, 1};
	CUtensorMapSwizzle sw;
	if (kmajor) { gdim[0] = (cuuint64_t)K; gdim[1] = (cuuint64_t)rows; box[0] = BK; box[1] = (cuuint32_t)box_rows; sw = CU_TENSOR_MAP_SWIZZLE_128B; }
	else        { gdim[0] = (cuuint64_t)rows; gdim[1] = (cuuint64_t)K; box[0] = 32; box[1] = BK; sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; }
	gdim[2] = (cuuint64_t)(batch > 0 ? batch : 1);
	box[2] = 1;
	gstride[0] = (cuuint64_t)ld * 4;
	gstride[1] = (batch > 1) ? (cuuint64_t)stride * 4 : gstride[0] * gdim[1];   // any legal value when there is one instance
	return cached_encode(map, 3, base, gdim, gstride, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// C (M lines of N fp32, pitch ldc) as the target of 32-row x 32-column box stores: dims {N, M, batch}, SWIZZLE_128B (the box's
// 128-byte rows are staged swizzled so the epilogue's per-row 16-byte shared-memory stores are conflict-free)
bool make_c_map(CUtensorMap *map, float *base, long long M, long long N, long long ldc, int batch, long long strideC)
{
	cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)(batch > 0 ? batch : 1)};
	cuuint64_t gstride[2] = {(cuuint64_t)ldc * 4, (batch > 1) ? (cuuint64_t)strideC * 4 : (cuuint64_t)ldc * 4 * (cuuint64_t)M};
	cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
	return cached_encode(map, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
}

unsigned *g_diag_host = nullptr, *g_diag_dev = nullptr;
unsigned *diag_dev()
{
	static std::once_flag once;
	std::call_once(once, [] {
		if (cudaHostAlloc(&g_diag_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
			for (int i = 0; i < 16; i++) g_diag_host[i] = 0;
			if (cudaHostGetDevicePointer(&g_diag_dev, g_diag_host, 0) != cudaSuccess) g_diag_dev = nullptr;
		}
	});
	return g_diag_dev;
}

// common tail of the GEMM and convolution launches: scheduler counters, attributes, cluster launch
template <int CG, bool CONV, bool TS = false>
cudaError_t launch_kernel(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const CUtensorMap &tmW, K1Params &P, long long nt, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	if (nt > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
	P.num_tiles = (int)nt;
	P.kc_blocks = (t.kc_blocks > 0 && t.kc_blocks < P.num_k_blocks) ? t.kc_blocks : P.num_k_blocks;
	P.split = t.split;
	P.flags = t.flags;
	P.group = TS && ((t.flags >> 24) & 31) ? ((t.flags >> 24) & 31) : 8;       // (bits 24-28 of the flags: tile-order experiments)
	// serpentine K where waves re-read panels from DRAM: operands well beyond the L2 and at least four tiles either way
	P.serpentine = TS && !CONV && !(t.flags & 524288) && P.tiles_m >= 4 && P.tiles_n >= 4 &&
	               4.0 * ((double)P.M * P.K + (double)P.K * P.N) > 96e6;
	P.diag = diag_dev();
	// Device-resident launch state is kept PER DEVICE (the single-process multi-GPU driver, sgemm_cuda_mgpu, launches this
	// kernel on every GPU of the box from one host thread): scheduler counters, profiling buffer, function attributes.
	constexpr int MAX_DEV = 32;
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
	// (sgemm_cuda_dev may be called from several host threads on their own streams: the one-time set-up below is serialised)
	static std::mutex init_mu;
	std::lock_guard<std::mutex> init_lock(init_mu);
	// dynamic-scheduler counters: a small pool so launches on different streams do not share a slot
	static unsigned *sched_pools[MAX_DEV] = {nullptr};
	static unsigned sched_claimed[MAX_DEV][64] = {{0}};     // host shadow: value each device counter will have after the launches queued so far
	static unsigned sched_next = 0;
	constexpr unsigned SCHED_POOL = 64;
	if (!sched_pools[dev]) {
		if (cudaMalloc(&sched_pools[dev], SCHED_POOL * sizeof(unsigned)) != cudaSuccess) return cudaErrorMemoryAllocation;
		cudaMemset(sched_pools[dev], 0, SCHED_POOL * sizeof(unsigned));
		cudaDeviceSynchronize();
	}
	const unsigned sched_slot = sched_next++ % SCHED_POOL;
	P.sched = sched_pools[dev] + sched_slot;
	P.sched_base = sched_claimed[dev][sched_slot];
	P.prof = nullptr;
	static long long *prof_devs[MAX_DEV] = {nullptr};
	const bool prof = !CONV && (t.flags & 32);
	if (prof) {
		if (!prof_devs[dev]) cudaMalloc(&prof_devs[dev], 64 * sizeof(long long));
		cudaMemsetAsync(prof_devs[dev], 0, 64 * sizeof(long long), stream);
		P.prof = prof_devs[dev];
	}
	long long *const prof_dev = prof_devs[dev];

	static bool attr_sets[MAX_DEV] = {false};      // per device, one array per <CG, CONV, TS> instantiation of this function
	bool &attr_set = attr_sets[dev];
	constexpr int smem_bytes = TS ? tsk::TS_SMEM_BYTES : SMEM_BYTES;
	if (!attr_set) {
		cudaError_t e = TS ? cudaFuncSetAttribute(k1ts_kernel<CG, false, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
		                   : cudaFuncSetAttribute(k1_3xtf32_kernel<CG, false, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
		if (e != cudaSuccess) return e;
		if (!CONV) {
			e = TS ? cudaFuncSetAttribute(k1ts_kernel<CG, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
			       : cudaFuncSetAttribute(k1_3xtf32_kernel<CG, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
			if (e != cudaSuccess) return e;
		}
		attr_set = true;
	}
	const int max_clusters = sm_count / CG;
	const int clusters = (int)(nt < max_clusters ? nt : max_clusters);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(clusters * CG));
	cfg.blockDim = dim3(NUM_THREADS);
	cfg.dynamicSmemBytes = smem_bytes;
	cfg.stream = stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	// TS kernel: programmatic dependent launch (the kernel waits with griddepcontrol.wait before it reads anything); flags bit 18 = off
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = (TS && !(t.flags & 262144)) ? 2 : 1;
	cudaError_t le = TS   ? (prof ? cudaLaunchKernelEx(&cfg, k1ts_kernel<CG, true, false>, tmA, tmB, tmC, tmW, P) : cudaLaunchKernelEx(&cfg, k1ts_kernel<CG, false, CONV>, tmA, tmB, tmC, tmW, P))
	               : prof ? cudaLaunchKernelEx(&cfg, k1_3xtf32_kernel<CG, true, false>, tmA, tmB, tmC, P)
	                      : cudaLaunchKernelEx(&cfg, k1_3xtf32_kernel<CG, false, CONV>, tmA, tmB, tmC, P);
	// a dynamically scheduled launch advances its counter by one claim per tile plus one failed claim per cluster (a launch that
	// the runtime rejected never ran; a kernel that trapped poisons the context, so no later launch can observe the slot)
	// (TS kernel, stream-K launch: the whole tiles are claimed dynamically too -- sk_full successful claims, one failed claim per cluster)
	if (le == cudaSuccess && (P.sk_q <= 0 || TS)) sched_claimed[dev][sched_slot] += (unsigned)(P.sk_q > 0 ? P.sk_full : P.num_tiles) + (unsigned)clusters;
	if (le == cudaSuccess && prof) {   // debug: per-role cycle breakdown of CTAs 0..3 on stderr
		long long h[64];
		if (cudaStreamSynchronize(stream) == cudaSuccess && cudaMemcpy(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess)
			for (int c = 0; c < 4; c++)
				fprintf(stderr, "k1prof cta%d producer: wait_empty %lld / %lld | mma: wait_xf %lld wait_tempty %lld / %lld | transform: wait_full %lld work %lld fence %lld / %lld | epilogue: wait_tfull %lld drain %lld store %lld / %lld\n",
				        c, h[16 * c + 0], h[16 * c + 1], h[16 * c + 2], h[16 * c + 3], h[16 * c + 4], h[16 * c + 5], h[16 * c + 6], h[16 * c + 7],
				        h[16 * c + 8], h[16 * c + 9], h[16 * c + 10], h[16 * c + 11], h[16 * c + 12]);
	}
	return le;
}

// Stream-K tail.  A persistent grid of `pairs` clusters finishes nt tiles in ceil(nt / pairs) rounds; when the last round is
// only partly filled (c3: 192 tiles on 74 pairs = 2.6 rounds, 4096^3: 3.5, anything smaller than the machine: < 1), its
// tiles are cut along K into equal chunk ranges, one per pair, whose partial sums meet in a workspace (decode_item,
// k1_tail_fixup_kernel).  Taken when it shortens the last round by at least 15 % and the launch by at least 12 %.
// The decision as pure host arithmetic (also exported through sgemm_cuda_k1_plan, so that the partition is checked without a GPU):
// nt tiles of nkb k-blocks on `pairs` CTA pairs (or single CTAs), promotion chunks of kc k-blocks.  Returns true when the tail is cut,
// with full = whole tiles, rem = tail tiles, nch = chunks per tile, q = chunks per range, ranges = number of ranges.
struct TailPlan { long long full, rem, q, ranges; int nch; };
bool plan_tail(long long nt, int nkb, int kc_eff, long long pairs, bool ts, int flags, TailPlan *tp)
{
	const int nch = (nkb + kc_eff - 1) / kc_eff;
	tp->full = nt; tp->rem = 0; tp->q = 0; tp->ranges = 0; tp->nch = nch;
	if ((flags & 2048) || nch < 2 || pairs <= 0 || nt % pairs == 0) return false;
	const long long full = nt / pairs * pairs, rem = nt - full;
	const long long q = (rem * nch + pairs - 1) / pairs, ranges = (rem * nch + q - 1) / q;
	// expected saving: (1 - q/nch) of one round out of ceil(nt / pairs); the fix-up pass and the tail's poorer L2 locality
	// (parts of one tile run at different k offsets) cost a few percent of a round, so small savings are not worth it
	const double saved_rounds = 1.0 - (double)q / nch, rounds = (double)((nt + pairs - 1) / pairs);
	// (SS kernel, measured: a modelled saving of 7-11 % of the launch came out as a 2-4 % loss, 13 % as a 10 % gain)
	// TS kernel [measured, profiles/r2n_sk_sweep.jsonl, r2p_sk_sweep.jsonl]: once at least one full round precedes the tail, the part
	// stores, the fix-up pass and the tail's colder start cost about as much as 30 k-blocks of a pair; below that the tail is a loss
	// (4096 x 3072 x 2048: 24 k-blocks saved, 219 vs 209 us), above it a gain (2560^3: 48 saved, 155 vs 175 us; 4096^3: 68, 494 vs 522)
	// ... and at least 2 % of a pair's whole work: at 8192^3 the tail saves 40 of 3543 k-blocks per pair, costs 0.36 GB of extra DRAM
	// traffic (parts, colder tail) and measures as nothing (3.66 vs 3.64 ms, profiles/r3a_traffic.csv)
	const long long saved_kb = (nch - q) * (long long)kc_eff, pair_kb = (nt * (long long)nkb + pairs - 1) / pairs;
	const bool worth = ts ? (full == 0 ? saved_rounds >= 0.15 : (saved_kb >= 32 && saved_kb * 50 >= pair_kb))
	                      : (saved_rounds >= 0.15 && saved_rounds / rounds >= 0.12);
	if (!((worth || ((flags & 131072) && saved_rounds > 0.0)) && rem * nch < 0x3fffffffLL)) return false;      // (bit 17: tail whenever it saves anything, A/B runs)
	tp->full = full; tp->rem = rem; tp->q = q; tp->ranges = ranges;
	return true;
}

template <int CG, bool CONV, bool TS = false>
cudaError_t launch_with_tail(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, K1Params &P, long long nt, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	const int kc_eff = (t.kc_blocks > 0 && t.kc_blocks < P.num_k_blocks) ? t.kc_blocks : P.num_k_blocks;
	const long long pairs = sm_count / CG;
	const int tile_m = 128 * CG, tile_n = 128 * CG;
	long long items = nt;
	float *ws = nullptr;
	TailPlan tp;
	if (plan_tail(nt, P.num_k_blocks, kc_eff, pairs, TS, t.flags, &tp)) {
		const size_t tile_bytes = (size_t)tile_m * tile_n * sizeof(float);
		if (cudaMallocAsync(reinterpret_cast<void **>(&ws), (size_t)(2 * tp.ranges) * tile_bytes, stream) == cudaSuccess) {
			P.sk_full = (int)tp.full; P.sk_rem = (int)tp.rem; P.sk_nch = tp.nch; P.sk_q = (int)tp.q; P.sk_ws = ws;
			items = tp.full + tp.ranges;
		} else { cudaGetLastError(); ws = nullptr; }
	}
	CUtensorMap tmW = tmA;
	if (TS && ws) {
		// the parts of the tail travel through the TMA unit as well: workspace = {tile_n columns, slots * tile_m rows}, 32 x 32 boxes
		cuuint64_t gdim[2] = {(cuuint64_t)tile_n, (cuuint64_t)(2 * (items - P.sk_full)) * tile_m};
		cuuint64_t gstride[1] = {(cuuint64_t)tile_n * 4};
		cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
		if (!cached_encode(&tmW, 2, ws, gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { cudaFreeAsync(ws, stream); return cudaErrorInvalidValue; }
	}
	cudaError_t e = launch_kernel<CG, CONV, TS>(tmA, tmB, tmC, tmW, P, items, t, stream, sm_count);
	if (ws) {
		if (e == cudaSuccess && !(t.flags & 65536)) {      // (bit 16: ablation, fix-up pass skipped)
			cudaLaunchConfig_t fc = {};
			fc.gridDim = dim3((unsigned)P.sk_rem, (unsigned)(tile_m / FIXUP_ROWS));
			fc.blockDim = dim3(256);
			fc.stream = stream;
			cudaLaunchAttribute fa[1];
			fa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
			fa[0].val.programmaticStreamSerializationAllowed = 1;
			fc.attrs = fa; fc.numAttrs = (t.flags & 262144) ? 0 : 1;
			e = cudaLaunchKernelEx(&fc, k1_tail_fixup_kernel<CG>, P);
		}
		cudaFreeAsync(ws, stream);
	}
	return e;
}

template <int CG>
cudaError_t launch_cg(const Problem &p, const K1Tuning &t_in, cudaStream_t stream, int sm_count)
{
	// The TS kernel (A operand in tensor memory, k1ts_kernel) is the production kernel; flags bit 15 (32768) selects the round-1 SS
	// kernel (both operands from shared memory) for A/B runs, and the RNA split experiment stays on it (truncation split only in TS).
	const bool ts = !(t_in.flags & 32768) && t_in.split == 0;
	K1Tuning t = t_in;
	if (ts && t.kc_blocks > 0) {            // the TS kernel hands over one 64-column slice every kc / NSL k-blocks: kc a multiple of NSL
		const int nsl = 2 * CG;
		t.kc_blocks = (t.kc_blocks + nsl - 1) / nsl * nsl;
	}
	CUtensorMap tmA, tmB;
	if (!make_operand_map(&tmA, p.A, p.M, p.K, p.lda, p.a_kmajor, p.batch, p.strideA)) return cudaErrorInvalidValue;
	// TS kernel: K-major B is loaded in 32-row groups (see k1ts_kernel)
	if (!make_operand_map(&tmB, p.B, p.N, p.K, p.ldb, p.b_kmajor, p.batch, p.strideB, ts ? 32 : ROWS)) return cudaErrorInvalidValue;
	K1Params P = {};
	P.M = p.M; P.N = p.N; P.K = p.K; P.alpha = p.alpha; P.beta = p.beta; P.C = p.C; P.ldc = p.ldc;
	P.bias = p.bias; P.slope = p.slope;
	P.a_kmajor = p.a_kmajor; P.b_kmajor = p.b_kmajor;
	const int tile_m = 128 * CG, tile_n = 128 * CG;
	P.tiles_m = (p.M + tile_m - 1) / tile_m;
	P.tiles_n = (p.N + tile_n - 1) / tile_n;
	P.tiles_per_batch = P.tiles_m * P.tiles_n;
	P.strideC = p.strideC;
	const long long nt = (long long)P.tiles_m * P.tiles_n * (p.batch > 0 ? p.batch : 1);
	P.num_k_blocks = (p.K + BK - 1) / BK;
	// 128-bit (and wider) accesses to rows of C: every row of every instance must start on a 16-byte boundary
	P.vecC = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 4 == 0 && (p.batch <= 1 || p.strideC % 4 == 0)) ? 1 : 0;
	// C as a TMA-store target: {N, M, batch} with 32 x 32 boxes (flags bit 13 = 8192 switches the TMA-store epilogue off)
	CUtensorMap tmC = tmA;
	P.tma_store = 0;
	if (P.vecC && !(t.flags & 8192) && (p.batch <= 1 || p.strideC % 4 == 0) && make_c_map(&tmC, p.C, p.M, p.N, p.ldc, p.batch, p.strideC)) P.tma_store = 1;

	if (ts) return launch_with_tail<CG, false, true>(tmA, tmB, tmC, P, nt, t, stream, sm_count);
	return launch_with_tail<CG, false>(tmA, tmB, tmC, P, nt, t, stream, sm_count);
}

// channels-last image as a 4-D tensor {c: cs, x: w, y: h, image: nimg}, box {32 c, 32 x, 1 y, 1}: one box = 32 output pixels of one
// output row x 32 input channels at one kernel position = 32 rows of 128 B, laid out like 32 rows of a dense K-major tile
bool make_image_map(CUtensorMap *map, const ConvProblem &c)
{
	cuuint64_t gdim[4] = {(cuuint64_t)c.cs, (cuuint64_t)c.w, (cuuint64_t)c.h, (cuuint64_t)c.nimg};
	cuuint64_t gstride[3] = {(cuuint64_t)c.cs * 4, (cuuint64_t)c.cs * c.w * 4, (cuuint64_t)c.cs * c.w * c.h * 4};
	// a strided convolution reads every stride-th pixel of the row: the box spans 32*stride pixels and the TMA element stride
	// picks 32 of them
	cuuint32_t box[4] = {32, (cuuint32_t)(32 * c.stride), 1, 1}, estr[4] = {1, (cuuint32_t)c.stride, 1, 1};
	return cached_encode(map, 4, c.in_hwc, gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

template <int CG>
cudaError_t launch_conv_cg(const ConvProblem &c, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	const int kk = c.k * c.k * c.ichp, npix = c.ho * c.wo, wp = (c.wo + 31) / 32 * 32;
	CUtensorMap tmA, tmB;
	if (!make_operand_map(&tmA, c.wgt_kkc, c.ch, kk, kk, true, 1, 0)) return cudaErrorInvalidValue;
	if (!make_image_map(&tmB, c)) return cudaErrorInvalidValue;
	K1Params P = {};
	P.M = c.ch; P.N = c.ho * wp; P.K = kk; P.alpha = 1.f; P.beta = 0.f; P.C = c.out; P.ldc = npix;
	P.bias = c.bias; P.slope = c.slope;
	P.a_kmajor = 1; P.b_kmajor = 1;
	const int tile_m = 128 * CG, tile_n = 128 * CG;
	P.tiles_m = (P.M + tile_m - 1) / tile_m;
	P.tiles_n = (P.N + tile_n - 1) / tile_n;
	P.tiles_per_batch = P.tiles_m * P.tiles_n;
	P.strideC = (long long)c.ch * npix;
	const long long nt = (long long)P.tiles_m * P.tiles_n * c.nimg;
	P.num_k_blocks = kk / BK;
	P.vecC = ((reinterpret_cast<uintptr_t>(c.out) & 15) == 0 && npix % 4 == 0) ? 1 : 0;
	P.cv_wp = wp; P.cv_wo = c.wo; P.cv_ho = c.ho; P.cv_k = c.k; P.cv_pad = c.pad; P.cv_cblocks = c.ichp / 32; P.cv_npix = npix; P.cv_stride = c.stride;
	// output [img][co][io][jo] as a TMA-store target {wo, ho, ch, img}, box {32 x, 1 y, 32 filters, 1}: needs 16-byte row pitch
	CUtensorMap tmC = tmA;
	P.tma_store = 0;
	if (P.vecC && c.wo % 4 == 0 && !(t.flags & 8192)) {
		cuuint64_t gdim[4] = {(cuuint64_t)c.wo, (cuuint64_t)c.ho, (cuuint64_t)c.ch, (cuuint64_t)c.nimg};
		cuuint64_t gstride[3] = {(cuuint64_t)c.wo * 4, (cuuint64_t)npix * 4, (cuuint64_t)c.ch * npix * 4};
		cuuint32_t box[4] = {32, 1, 32, 1}, estr[4] = {1, 1, 1, 1};
		if (cached_encode(&tmC, 4, c.out, gdim, gstride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE)) P.tma_store = 1;
	}
	// production: the TS kernel (flags bit 15 selects the round-1 SS kernel for A/B runs); kc a multiple of the slice count
	if (!(t.flags & 32768) && t.split == 0) {
		K1Tuning tt = t;
		if (tt.kc_blocks > 0) { const int nsl = 2 * CG; tt.kc_blocks = (tt.kc_blocks + nsl - 1) / nsl * nsl; }
		return launch_with_tail<CG, true, true>(tmA, tmB, tmC, P, nt, tt, stream, sm_count);
	}
	return launch_with_tail<CG, true>(tmA, tmB, tmC, P, nt, t, stream, sm_count);
}

// planar [img][c][y][x] -> channels-last [img][y][x][cs] (cs = ich rounded up to 4, the pad channels zero): one image-sized HBM pass
// through a 32 x 32 shared-memory tile so that both sides are coalesced.  blockIdx.z = img * h + y.
__global__ void __launch_bounds__(256)
chw_to_hwc_kernel(const float *__restrict__ in, int ich, int h, int w, int cs, float *__restrict__ out)
{
	__shared__ float tile[32][33];
	const int img = blockIdx.z / h, y = blockIdx.z - img * h;
	const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const float *src = in + ((long long)img * ich * h + y) * w;         // + c * h * w + x
#pragma unroll
	for (int r = ty; r < 32; r += 8) {
		const int c = c0 + r, x = x0 + tx;
		tile[r][tx] = (c < ich && x < w) ? __ldg(src + (long long)c * h * w + x) : 0.f;
	}
	__syncthreads();
	float *dst = out + (((long long)img * h + y) * w) * cs;             // + x * cs + c
#pragma unroll
	for (int r = ty; r < 32; r += 8) {
		const int x = x0 + r, c = c0 + tx;
		if (x < w && c < cs) dst[(long long)x * cs + c] = tile[tx][r];
	}
}

// The same pass over FLAT pixels (p = y * w + x is contiguous in a planar image, so [c][p] -> [p][c] is a plain 2-D transpose per
// image and a row width like 56 leaves no ragged x tile), 32 channels x 128 pixels per block: a warp reads 512 contiguous bytes of
// one channel with one 128-bit load per lane (four in flight per thread: the 32 x 32 kernel above keeps 32 KiB in flight per SM and
// runs at 3.7 TB/s, Little's law asks for ~ 44 KiB), and writes four pixels x 128 bytes per instruction.  Needs h * w % 4 == 0 and
// 16-byte aligned bases.  blockIdx = (pixel tile, channel tile, image).
__global__ void __launch_bounds__(256)
chw_to_hwc_flat_kernel(const float *__restrict__ in, int ich, int npix, int cs, float *__restrict__ out)
{
	__shared__ float tile[32][129];          // [channel][pixel]; odd row stride: the column reads of the write phase hit 32 different banks
	const int p0 = blockIdx.x * 128, c0 = blockIdx.y * 32, img = blockIdx.z;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const float *src = in + (long long)img * ich * npix;
	float4 v[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int c = c0 + warp + 8 * i, p = p0 + 4 * lane;
		v[i] = (c < ich && p < npix) ? __ldg(reinterpret_cast<const float4 *>(src + (long long)c * npix + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		float *t = &tile[warp + 8 * i][4 * lane];
		t[0] = v[i].x; t[1] = v[i].y; t[2] = v[i].z; t[3] = v[i].w;
	}
	__syncthreads();
	float *dst = out + (long long)img * npix * cs;
	const int cq = 4 * (lane & 7), pl = lane >> 3;     // this lane's four channels, its pixel within a group of four
	// 128 pixels x 32 channels = 1024 float4; 256 threads x 4: pixel = 4 * (8 i + warp) + pl
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const int pt = 4 * (8 * i + warp) + pl, p = p0 + pt, c = c0 + cq;
		if (p < npix && c < cs)
			*reinterpret_cast<float4 *>(dst + (long long)p * cs + c) = make_float4(tile[cq][pt], tile[cq + 1][pt], tile[cq + 2][pt], tile[cq + 3][pt]);
	}
}

// dst[co][(ki*k + kj)*ichp + c] = w[co][c][ki][kj], zero for ich <= c < ichp
__global__ void conv_weight_repack_kernel(const float *__restrict__ w, int ch, int ich, int k, int ichp, float *__restrict__ dst)
{
	const long long total = (long long)ch * k * k * ichp;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int c = (int)(i % ichp);
		const long long r = i / ichp;
		const int kpos = (int)(r % (k * k)), co = (int)(r / (k * k));
		dst[i] = c < ich ? __ldg(w + ((long long)co * ich + c) * k * k + kpos) : 0.f;
	}
}

} // namespace

const unsigned *k1_diag_host() { return g_diag_host; }

bool k1_eligible(const Problem &p, const char **why)
{
	const char *w = nullptr;
	if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) w = "A or B not 16-byte aligned (TMA base rule)";
	else if (p.lda % 4 || p.ldb % 4) w = "lda or ldb not a multiple of 4 (TMA global stride must be a multiple of 16 bytes)";
	else if (p.batch > 1 && (p.strideA % 4 || p.strideB % 4)) w = "batch strides of A or B not multiples of 4 (TMA global stride rule)";
	else if (p.M < 1 || p.N < 1 || p.K < 1) w = "empty problem";
	else if (!encode_fn()) w = "cuTensorMapEncodeTiled unavailable";
	if (why) *why = w;
	return w == nullptr;
}

// CTA pairing of a dense problem (pure host arithmetic): 2 = 256 x 256 tiles on CTA pairs, 1 = 128 x 128 tiles on single CTAs
int choose_cta_group(int M, int N, int K, int batch, const K1Tuning &t, int sm_count)
{
	struct { int M, N, K, batch; } p = {M, N, K, batch};
	int cg = t.cta_group;
	if (cg != 1 && cg != 2) {
		// auto: 2-CTA pairs (256x256 tiles, half the shared-memory operand traffic per flop) once there are enough
		// such tiles to occupy ~3/4 of the SM pairs; below that 128x128 single-CTA tiles spread the problem over more
		// SMs (measured cross-over between 1536^3 = 36 pair tiles and 2048^3 = 64, profiles/r1_sizes.txt)
		const long long pair_tiles = (long long)((p.M + 255) / 256) * ((p.N + 255) / 256) * (p.batch > 0 ? p.batch : 1);
		cg = (pair_tiles * 8 >= (long long)(sm_count / 2) * 6) ? 2 : 1;
		// with the stream-K tail, fewer pair tiles still fill the machine when K is long enough to give every pair >= 4
		// promotion chunks (1024 x 1024 x 8192: 109 vs 134 us; 1536^3: 58 vs 62 us; profiles/r1_sizes.txt)
		const int nkb = (p.K + BK - 1) / BK, kc_eff = (t.kc_blocks > 0 && t.kc_blocks < nkb) ? t.kc_blocks : nkb;
		if (cg == 1 && !(t.flags & 2048) && p.batch <= 1 && pair_tiles * ((nkb + kc_eff - 1) / kc_eff) >= 4LL * (sm_count / 2)) cg = 2;
		// a side of at most 128 fits one single-CTA tile: pairing would only multiply by zero-filled rows or columns
		// (200704 x 128 x 1152: 0.42 vs 0.50 ms; 8192 x 64 x 8192: 0.14 vs 0.17 ms; profiles/r1_skinny_k1_vs_k2.jsonl)
		if (p.M <= 128 || p.N <= 128) cg = 1;
	}
	return cg;
}

cudaError_t launch_k1_3xtf32(const Problem &p, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	if (choose_cta_group(p.M, p.N, p.K, p.batch, t, sm_count) == 1) return launch_cg<1>(p, t, stream, sm_count);
	return launch_cg<2>(p, t, stream, sm_count);
}

// The schedule of a dense K1 launch as numbers (sgemm_cuda_k1_plan): out[0..11] = cta_group, tile_m, tile_n, tiles_m, tiles_n,
// k-blocks per tile, promotion interval in k-blocks, whole tiles, tail tiles, chunks per tile, chunks per tail range, work items.
void k1_plan(int M, int N, int K, int batch, const K1Tuning &t_in, int sm_count, int *out)
{
	const bool ts = !(t_in.flags & 32768) && t_in.split == 0;
	const int cg = choose_cta_group(M, N, K, batch, t_in, sm_count);
	K1Tuning t = t_in;
	if (ts && t.kc_blocks > 0) { const int nsl = 2 * cg; t.kc_blocks = (t.kc_blocks + nsl - 1) / nsl * nsl; }
	const int tile = 128 * cg, tiles_m = (M + tile - 1) / tile, tiles_n = (N + tile - 1) / tile, nkb = (K + BK - 1) / BK;
	const long long nt = (long long)tiles_m * tiles_n * (batch > 0 ? batch : 1);
	const int kc_eff = (t.kc_blocks > 0 && t.kc_blocks < nkb) ? t.kc_blocks : nkb;
	TailPlan tp;
	const bool tail = plan_tail(nt, nkb, kc_eff, sm_count / cg, ts, t.flags, &tp);
	const long long v[12] = {cg, tile, tile, tiles_m, tiles_n, nkb, kc_eff, tail ? tp.full : nt, tail ? tp.rem : 0, tp.nch, tail ? tp.q : 0, tail ? tp.full + tp.ranges : nt};
	for (int i = 0; i < 12; i++) out[i] = (int)v[i];
}
// one segment of a work item of that schedule (decode_item, the function every role of the kernel decodes items with)
void k1_plan_item(int item, int h, int sk_full, int sk_rem, int sk_nch, int sk_q, int kc, int nkb, int *out)
{
	const Item it = decode_item(item, h, sk_full, sk_rem, sk_nch, sk_q, kc, nkb);
	out[0] = it.tile; out[1] = it.kb0; out[2] = it.kb1; out[3] = it.slot;
}

cudaError_t launch_conv_weight_repack(const float *w, int ch, int ich, int k, int ichp, float *dst, cudaStream_t stream)
{
	const long long total = (long long)ch * k * k * ichp;
	if (total <= 0) return cudaSuccess;
	long long blocks = (total + 255) / 256;
	if (blocks > 148LL * 16) blocks = 148LL * 16;
	conv_weight_repack_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, ch, ich, k, ichp, dst);
	return cudaGetLastError();
}

cudaError_t launch_chw_to_hwc(const float *in, int nimg, int ich, int h, int w, int cs, float *out, cudaStream_t stream)
{
	if (h < 1 || h > 65535) return cudaErrorInvalidConfiguration;
	const long long npix = (long long)h * w;
	static const bool old_pass = getenv("UGEMM_CONV_STAGE_OLD") != nullptr;      // (A/B runs)
	if (!old_pass && npix % 4 == 0 && npix < (1LL << 30) && nimg <= 65535 && !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15)) {
		dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((cs + 31) / 32), (unsigned)nimg);
		chw_to_hwc_flat_kernel<<<grid, 256, 0, stream>>>(in, ich, (int)npix, cs, out);
		return cudaGetLastError();
	}
	const int per_launch = 65535 / h;                 // grid.z = images x rows is limited to 65535
	for (int i0 = 0; i0 < nimg; i0 += per_launch) {
		const int n = nimg - i0 < per_launch ? nimg - i0 : per_launch;
		dim3 grid((unsigned)((w + 31) / 32), (unsigned)((cs + 31) / 32), (unsigned)(n * h));
		chw_to_hwc_kernel<<<grid, 256, 0, stream>>>(in + (size_t)i0 * ich * h * w, ich, h, w, cs, out + (size_t)i0 * h * w * cs);
	}
	return cudaGetLastError();
}

cudaError_t launch_k1_conv(const ConvProblem &c, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	if (!encode_fn()) return cudaErrorNotSupported;
	int cg = t.cta_group;
	if (cg != 1 && cg != 2) {
		const long long wp = (c.wo + 31) / 32 * 32;
		const long long pair_tiles = (long long)((c.ch + 255) / 256) * ((c.ho * wp + 255) / 256) * c.nimg;
		cg = (pair_tiles * 8 >= (long long)(sm_count / 2) * 6) ? 2 : 1;
		if (c.ch <= 128 || (long long)c.ho * wp <= 128) cg = 1;      // one single-CTA tile covers that side (see launch_k1_3xtf32)
	}
	if (cg == 1) return launch_conv_cg<1>(c, t, stream, sm_count);
	return launch_conv_cg<2>(c, t, stream, sm_count);
}

cudaError_t launch_probe_tf32(const float *dA, const float *dB, float *dD, int ksteps, cudaStream_t stream)
{
	if (ksteps < 1 || ksteps > 4) return cudaErrorInvalidValue;
	probe_tf32_kernel<<<1, 128, 0, stream>>>(dA, dB, dD, ksteps, diag_dev());
	return cudaGetLastError();
}

} // namespace ugemm
