// k1_tcgen05.cu -- K1: 3xTF32 error-compensated SGEMM on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
// C = alpha * op(A) op(B) + beta * C with fp32-class accuracy from TF32 tensor-core products:
//     a = a_big + a_small,  a_big = tf32(a) (top 19 bits),  a_small = a - a_big  (exact in fp32)
//     a*b ~= a_small*b_big + a_big*b_small + a_big*b_big        (a_small*b_small ~ 2^-22 dropped)
// Three tcgen05.mma per k-step, credited as 2*M*N*K flop (effective roofline = dense TF32 peak / 3).
//
// Replaces the reference's blocked inner loops (sgemm_avx256.h:20-390, gemm_cpu.h:96-282, the OpenCL
// gemm_fast / gemm_rnn kernels sgemm_ocl.h:444-538, sgemm_ocl2.h:17-90) for TMA-eligible problems.
//
// Structure (one CTA per SM, optionally paired as a 2-CTA cluster issuing cta_group::2 MMAs; 640 threads):
//   warp 0      TMA producer: raw fp32 tiles of op(A) (128 rows) and op(B) (128 rows) per k-block of 32,
//               128B-swizzled, into a 3-stage shared-memory ring.  K-major operands: one {32 k x 128 row}
//               box, SWIZZLE_128B.  MN-major operands (transA=='T' / transB=='N'): four {32 mn x 32 k}
//               boxes, SWIZZLE_128B_ATOM_32B (the only MN-major layout tcgen05 accepts for 32-bit types).
//               Ragged M/N/K edges are zero-filled by TMA out-of-bounds handling.  CONV instantiation: the B
//               boxes are gathered from a channels-last image with 4-D coordinates (implicit im2col).
//   warp 1      MMA issuer (leader CTA only): per k-step of 8: small*big, big*small, big*big into the TMEM
//               accumulator; tcgen05.commit releases the stage; accumulators are double-buffered in TMEM.
//   warp 2      work scheduler (leader CTA only): claims tile indices from a global atomic counter -- or, for a
//               stream-K launch, walks the static schedule -- and publishes them to every role of both CTAs
//               through a 4-deep shared-memory ring.
//   warps 4-11  transform (two warpgroups sharing every stage): read each landed stage, write the "small"
//               operand copies (same swizzled layout, so the transform is a flat element-wise pass),
//               fence.proxy.async, signal.
//   warps 12-19 epilogue: tcgen05.ld the accumulator, promote partial sums every kc_blocks k-blocks into fp32
//               registers with round-to-nearest adds, then fused alpha/beta(/bias/LeakyReLU) and direct global
//               stores (row per thread, 128 B contiguous per 32-column group) -- or, for a stream-K part, raw
//               partial sums into the workspace that k1_tail_fixup_kernel adds up.
//   Persistent: tiles are handed out in an L2-friendly grouped order (8 m-tiles share an n sweep).
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <cstring>
#include <mutex>
#include <type_traits>

namespace ugemm {

namespace {

using namespace ptx;

constexpr int BK = 32;                        // fp32 elements per k-block = one 128-byte swizzle line
constexpr int ROWS = 128;                     // rows of op(A) / rows of op(B) staged per CTA per k-block
constexpr int OPER_BYTES = ROWS * BK * 4;     // 16 KiB
constexpr int RAW_BYTES = 2 * OPER_BYTES;     // A raw | B raw
constexpr int STAGE_BYTES = 2 * RAW_BYTES;    // A raw | B raw | A small | B small = 64 KiB
constexpr int STAGES = 3;
constexpr int NUM_THREADS = 640;              // 20 warps, see role map above
constexpr int XF_GROUPS = 2;                  // transform warpgroups
constexpr bool XF_SPLIT_STAGE = true;         // true: both groups share every stage (half each); false: groups alternate k-blocks
constexpr int BAR_BYTES = 256;
constexpr int SCHED_SLOTS = 4;               // depth of the dynamic tile-index ring
constexpr int CSTAGE_BYTES = 32 * 32 * 4;    // per epilogue warp: one 32-row x 32-column fp32 box staged for a TMA store
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 8 * CSTAGE_BYTES + 1024; // ring | barrier block (padded to 1 KiB) | C staging | slack for 1024-B alignment
static_assert(BAR_BYTES <= 1024 && SMEM_BYTES <= 232448, "shared-memory budget of one CTA (227 KiB)");
constexpr long long WATCHDOG_CYCLES = 6000000000LL;

struct K1Params {
	int M, N, K;
	float alpha, beta;
	float *C;
	long long ldc;
	const float *bias;
	float slope;
	long long strideC;          // elements between batch instances of C
	int tiles_per_batch;        // tiles_m * tiles_n; tile index = instance * tiles_per_batch + tile within the instance
	int a_kmajor, b_kmajor;
	int tiles_m, tiles_n, num_tiles;
	int num_k_blocks, kc_blocks, split, vecC, flags;
	int tma_store;              // epilogue writes C through 32 x 32 TMA box stores (tmC valid: C 16-byte aligned, ldc % 4 == 0)
	// stream-K tail (sk_q > 0): work items [0, sk_full) are whole tiles; the remaining sk_rem tiles are cut into chunk ranges of
	// sk_q promotion chunks (kc_blocks k-blocks each, sk_nch per tile), two items per range (a range may straddle one tile
	// boundary); their raw partial sums go to sk_ws[slot][tile_m x tile_n] and k1_tail_fixup_kernel adds them up in range order
	int sk_full, sk_rem, sk_nch, sk_q;
	float *sk_ws;
	// implicit-GEMM convolution (CONV instantiation): padded output width (multiple of 32), output width / height,
	// kernel size, padding, 32-channel blocks per kernel position, pixels per output plane
	int cv_wp, cv_wo, cv_ho, cv_k, cv_pad, cv_cblocks, cv_npix, cv_stride;
	int group;                  // m-tiles that share an n sweep in the tile order (decode_tile)
	int serpentine;             // TS kernel: every other wave of tiles walks K downwards (see the producer)
	unsigned *diag;
	// dynamic scheduler: *sched is a device counter that only ever grows; a launch claims the values [sched_base, sched_base +
	// num_tiles + clusters) (every cluster makes exactly one claim past the end), so the host knows the base of the next launch
	// on this slot without any reset on the device (nothing to leave dirty, nothing for the last cluster to re-arm)
	unsigned *sched;
	unsigned sched_base;
	long long *prof;   // flags & 32: per-role cycle counters of the first 4 CTAs (16 slots each), debug only
};

__device__ __forceinline__ void watchdog_fail(unsigned *diag, int code, uint32_t parity)
{
	if (diag) {
		diag[1] = blockIdx.x; diag[2] = threadIdx.x; diag[3] = parity; diag[0] = (unsigned)code;
		__threadfence_system();
	}
	__trap();
}
// spin on an mbarrier phase with a watchdog so that a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned *diag, int code)
{
	if (mbar_try_wait(bar, parity)) return;
	const long long t0 = clock64();
	while (!mbar_try_wait(bar, parity))
		if (clock64() - t0 > WATCHDOG_CYCLES) watchdog_fail(diag, code, parity);
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, unsigned *diag, int code)
{
	if (mbar_try_wait_cluster(bar, parity)) return;
	const long long t0 = clock64();
	while (!mbar_try_wait_cluster(bar, parity))
		if (clock64() - t0 > WATCHDOG_CYCLES) watchdog_fail(diag, code, parity);
}

// grouped tile order: 8 consecutive m-tiles share an n sweep so a wave's A and B panels stay in L2
__device__ __forceinline__ void decode_tile(int tile, int tiles_m, int tiles_n, int &tm, int &tn, int GROUP = 8)
{
	const int per_group = GROUP * tiles_n;
	const int group = tile / per_group;
	const int first_m = group * GROUP;
	const int gsize = min(tiles_m - first_m, GROUP);
	const int r = tile - group * per_group;
	tm = first_m + r % gsize;
	tn = r / gsize;
}

// A work item of the dynamic scheduler -> one or two segments (tile, k-block range, workspace slot).  Items below sk_full
// are whole tiles (one segment, slot < 0: normal epilogue into C).  Item sk_full + r is chunk range r = [r*q, (r+1)*q) of the
// tail's sk_rem * sk_nch chunks; a range may straddle one tile boundary, so it has up to two segments: h = 0 inside the
// tile it starts in, h = 1 (possibly absent) in the next tile.  One pair processes a whole range, so the tail is balanced:
// every pair claims one range of q chunks.  Every role of the kernel decodes items with this one function, so they all agree.
struct Item { int tile, kb0, kb1, slot; };      // kb1 <= kb0: no such segment
__device__ __host__ __forceinline__ Item decode_item(int item, int h, int sk_full, int sk_rem, int sk_nch, int sk_q, int kc, int nkb)
{
	Item it;
	if (sk_q <= 0 || item < sk_full) { it.tile = item; it.kb0 = 0; it.kb1 = h == 0 ? nkb : 0; it.slot = -1; return it; }
	const int r = item - sk_full;
	const int total = sk_rem * sk_nch;
	const int lo = r * sk_q, hi = lo + sk_q < total ? lo + sk_q : total;
	const int ta = lo / sk_nch, bnd = (ta + 1) * sk_nch;
	const int tr = ta + h, c0 = h ? bnd : lo, c1 = h ? hi : (hi < bnd ? hi : bnd);
	it.tile = sk_full + tr;
	it.slot = 2 * r + h;
	it.kb0 = (c0 - tr * sk_nch) * kc;
	it.kb1 = (c1 - tr * sk_nch) * kc < nkb ? (c1 - tr * sk_nch) * kc : nkb;
	if (c1 <= c0) { it.kb0 = it.kb1 = 0; }
	return it;
}

// Arrive on a barrier that lives in the LEADER CTA of the pair, from either CTA, without a cluster-scope fence (see
// ptx.cuh: mbar_arrive_remote): the leader arrives locally, the peer through the cluster address.  `heavy` (UGEMM_K1_FLAGS bit 14,
// A/B runs) restores the round-1 form, a .release.cluster arrive from both CTAs.
template <int CG>
__device__ __forceinline__ void arrive_on_leader(uint32_t bar, uint32_t cta_rank, bool heavy)
{
	if (CG == 1) { mbar_arrive(bar); return; }
	if (heavy) mbar_arrive_cluster(bar, 0);
	else if (cta_rank == 0) mbar_arrive(bar);
	else mbar_arrive_remote(bar, 0);
}

// ---- dynamic tile scheduler ------------------------------------------------------------------------------------
// One thread per cluster (leader CTA, warp 2) claims tile indices from a global atomic counter and publishes them
// through a 4-deep shared-memory ring to every role of both CTAs; a CTA pair that starts late (SMs busy with another
// kernel, e.g. NCCL) simply claims fewer tiles.  sched_full[slot] (count 1, one per CTA) / sched_empty[slot] (leader
// only; one arrival per consuming role) are mbarriers; a negative index ends the kernel.
template <int CG>
__device__ __forceinline__ int next_tile(uint32_t bar_base, int &n, bool warp_collective, int lane, unsigned *diag, uint32_t cta_rank, bool heavy)
{
	const int slot = n & (SCHED_SLOTS - 1);
	const uint32_t ph = (n / SCHED_SLOTS) & 1;
	n++;
	const uint32_t full = bar_base + 8u * (14 + slot), empty = bar_base + 8u * (14 + SCHED_SLOTS + slot);
	if (CG == 2) mbar_wait_cluster(full, ph, diag, 6); else mbar_wait(full, ph, diag, 6);
	int tile;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tile) : "r"(bar_base + 8u * (14 + 2 * SCHED_SLOTS) + 4u * slot) : "memory");
	if (warp_collective) __syncwarp();
	if (!warp_collective || lane == 0) {
		arrive_on_leader<CG>(empty, cta_rank, heavy);
	}
	return tile;
}

template <bool PROF> __device__ __forceinline__ long long tick() { return PROF ? clock64() : 0LL; }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float small_part(float x, float big)
{
	return (__float_as_uint(x) & 0x7FFFFFFFu) == 0x7F800000u ? 0.f : x - big;
}
__device__ __forceinline__ float tf32_rna(float x)
{
	uint32_t r;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
	return __uint_as_float(r);
}

// ---- epilogue pieces shared by the SS kernel (k1_3xtf32_kernel) and the TS kernel (k1ts_kernel) -------------------------------
// An epilogue thread (lane quarter q, column half h) owns row `row` of the tile and NG = 2 * CG groups of 32 accumulator
// columns.  Tile-relative first column of group g: SS kernel -- the thread's half of the tile is contiguous, h * BN/2 + 32 g;
// TS kernel -- group g is the thread's half of 64-column accumulator slice g, 64 g + 32 h.
template <int CG, bool TS>
__device__ __forceinline__ int group_col(int h, int g) { return TS ? g * 64 + h * 32 : h * (64 * CG) + g * 32; }

// one group's running sums at the start of a tile (TS kernel: groups are re-armed one by one while the previous tile is stored)
// (TS kernel: the loaded values go into the accumulator registers UNTOUCHED -- no arithmetic on them, so nothing waits for the loads
// until the first promotion adds into them, k-blocks later; the beta/alpha weighting is carried by the promotion instead, see there)
template <int CG, bool TS>
__device__ __forceinline__ void epi_init_group(float (&a)[32], int g, const K1Params &P, bool from_c, const float *crow, int tn, int h)
{
	constexpr int BN = 128 * CG;
	if (from_c) {
		const long long col0 = (long long)tn * BN + group_col<CG, TS>(h, g);
		if (P.vecC && col0 + 31 < P.N) {
#pragma unroll
			for (int i = 0; i < 32; i += 4) {
				const float4 cv = *reinterpret_cast<const float4 *>(crow + col0 + i);
				a[i + 0] = cv.x; a[i + 1] = cv.y; a[i + 2] = cv.z; a[i + 3] = cv.w;
			}
		} else {
#pragma unroll
			for (int i = 0; i < 32; i++) a[i] = (col0 + i < P.N) ? crow[col0 + i] : 0.f;
		}
	} else {
#pragma unroll
		for (int i = 0; i < 32; i++) a[i] = 0.f;
	}
}

// running sums at the start of a tile: (beta/alpha) * C when the old C can be folded in up front, else 0
template <int CG, bool CONV, bool TS>
__device__ __forceinline__ void epi_init_acc(float (&acc)[2 * CG][32], const K1Params &P, const Item &wi, bool preload_c, float bs, long long row,
                                             const float *crow, int tn, int h)
{
	constexpr int BN = 128 * CG, NG = 2 * CG;
	// beta != 0: the old C is folded in UP FRONT -- the running sums start at (beta/alpha)*C, loaded while the
	// tile's first MMAs run and the epilogue warps would idle anyway -- so the tile end is store-only and
	// never stalls the accumulator hand-over on a global-load round trip.
	if (!CONV && preload_c && wi.slot < 0 && row < P.M) {
#pragma unroll
		for (int g = 0; g < NG; g++) {
			const long long col0 = (long long)tn * BN + group_col<CG, TS>(h, g);
			if (P.vecC && col0 + 31 < P.N) {
#pragma unroll
				for (int i = 0; i < 32; i += 4) {
					const float4 cv = *reinterpret_cast<const float4 *>(crow + col0 + i);
					acc[g][i + 0] = bs * cv.x; acc[g][i + 1] = bs * cv.y; acc[g][i + 2] = bs * cv.z; acc[g][i + 3] = bs * cv.w;
				}
			} else {
#pragma unroll
				for (int i = 0; i < 32; i++) acc[g][i] = (col0 + i < P.N) ? bs * crow[col0 + i] : 0.f;
			}
		}
	} else {
#pragma unroll
		for (int g = 0; g < NG; g++)
#pragma unroll
			for (int i = 0; i < 32; i++) acc[g][i] = 0.f;
	}
}

// tile end: stream-K part -> raw partial sums to the workspace; whole tile -> fused alpha/beta(/bias/LeakyReLU) and the store
// `after(g)` is called once per group, as soon as acc[g] has been consumed (staged for its TMA store / stored): the TS kernel
// re-arms the group for the next tile there and takes early hand-overs, so that the MMAs never wait for a tile store.
struct NoHook { __device__ __forceinline__ void operator()(int) const {} };
template <int CG, bool CONV, bool TS, class After = NoHook>
__device__ __forceinline__ void epi_store_tile(float (&acc)[2 * CG][32], const K1Params &P, const CUtensorMap *tmCp, const Item &wi, bool preload_c, float alpha,
                                               long long row, float *crow, int tm, int tn, int inst, int q, int h, int e, int lane, uint32_t cta_rank, uint32_t bar_base,
                                               After after = After(), const CUtensorMap *tmWp = nullptr)
{
	constexpr int BN = 128 * CG, UMMA_M = 128 * CG, NG = 2 * CG;
	// fused alpha/beta + store; ld padding and ragged edges are never written
	const float beta = preload_c ? 0.f : P.beta;   // already folded into acc when preloaded
	const bool part = wi.slot >= 0;                // stream-K part: raw partial sums to the workspace tile of this item
	const bool part_tma = part && tmWp != nullptr;
	if (part_tma || (!part && P.tma_store && beta == 0.f)) {
		// TMA-store epilogue: each warp stages one 32-row x 32-column box at a time in shared memory (128B-swizzled, so a
		// thread's eight 16-byte stores of its row are conflict-free) and hands it to the TMA unit, which writes whole
		// 128-byte lines and clips the box at the matrix edge -- instead of 32 row-strided 16-byte stores per instruction.
		// A stream-K part takes the same way into the workspace, a {BN, slots * UMMA_M} tensor whose tile `slot` starts at row
		// slot * UMMA_M: raw sums, no alpha, no bias.  (One loop for both, so that the `after` hook is expanded once per group.)
		const float slope = P.slope;
		const bool post = !part && (P.bias != nullptr || slope != 1.f);
		const float bm = (post && P.bias && row < P.M) ? __ldg(P.bias + row) : 0.f;
		const float scale = part ? 1.f : alpha;
		auto act = [&](float x) { x += bm; return x > 0.f ? x : x * slope; };
		const uint32_t cst = bar_base + 1024u + (uint32_t)e * CSTAGE_BYTES;
		const int row0 = (part ? wi.slot * UMMA_M : tm * UMMA_M) + (int)cta_rank * ROWS + q * 32;
		const bool skip = (P.flags & 16) != 0;         // ablation: nothing is stored
#pragma unroll
		for (int g = 0; g < NG; g++) {
			const int col0 = (part ? 0 : tn * BN) + group_col<CG, TS>(h, g);
			// CONV: the 32 columns are one output-row segment (io, jo0 .. jo0+31) of the padded column index
			const int io = (CONV && !part) ? col0 / P.cv_wp : 0, jo0 = (CONV && !part) ? col0 - io * P.cv_wp : 0;
			// warp-uniform: the whole box lies outside C
			const bool outside = !part && (row0 >= P.M || col0 >= P.N || (CONV && (io >= P.cv_ho || jo0 >= P.cv_wo)));
			if (!outside && !skip) {
				if (lane == 0) bulk_wait_group_read0();             // this warp's previous box has left shared memory
				__syncwarp();
#pragma unroll
				for (int i = 0; i < 32; i += 4) {
					float4 o;
					o.x = scale * acc[g][i + 0]; o.y = scale * acc[g][i + 1]; o.z = scale * acc[g][i + 2]; o.w = scale * acc[g][i + 3];
					if (post) { o.x = act(o.x); o.y = act(o.y); o.z = act(o.z); o.w = act(o.w); }
					sts128(cst + (uint32_t)lane * 128u + (uint32_t)(((i >> 2) ^ (lane & 7)) << 4), o);
				}
				fence_proxy_async_smem();
				__syncwarp();
				if (lane == 0) {
					if (part) tma_store_2d(tmWp, cst, col0, row0);
					else if (CONV) tma_store_4d(tmCp, cst, jo0, io, row0, inst);     // clipped at the output width and at the filter count
					else tma_store_3d(tmCp, cst, col0, row0, inst);
					bulk_commit_group();
				}
			}
			after(g);
		}
		return;
	}
	if (part) {
		// (SS kernel) stream-K part with plain stores: tile-local layout, UMMA_M x BN floats
		float *wrow = P.sk_ws + (long long)wi.slot * (UMMA_M * BN) + (long long)((int)cta_rank * ROWS + q * 32 + lane) * BN;
#pragma unroll
		for (int g = 0; g < NG; g++)
#pragma unroll
			for (int i = 0; i < 32; i += 4)
				*reinterpret_cast<float4 *>(wrow + group_col<CG, TS>(h, g) + i) = make_float4(acc[g][i], acc[g][i + 1], acc[g][i + 2], acc[g][i + 3]);
	} else
	if (row < P.M && !(P.flags & 16)) {
		const float slope = P.slope;
		const bool post = P.bias != nullptr || slope != 1.f;   // bias[row] + LeakyReLU (convolution callers)
		const float bm = P.bias ? __ldg(P.bias + row) : 0.f;
		auto act = [&](float x) { x += bm; return x > 0.f ? x : x * slope; };
		if (CONV) {
			// each 32-column group is one output-row segment: map it back from the padded column index
#pragma unroll
			for (int g = 0; g < NG; g++) {
				const int n0 = tn * BN + group_col<CG, TS>(h, g);
				const int io = n0 / P.cv_wp, jo0 = n0 - io * P.cv_wp;
				const int valid = io < P.cv_ho ? P.cv_wo - jo0 : 0;      // columns of this group that exist (may be <= 0 or >= 32)
				const int off = io * P.cv_wo + jo0;
				float *dst = crow + off;
				const bool vec = P.vecC && (off & 3) == 0;
#pragma unroll
				for (int i = 0; i < 32; i += 4) {
					float4 o;
					o.x = alpha * acc[g][i + 0]; o.y = alpha * acc[g][i + 1]; o.z = alpha * acc[g][i + 2]; o.w = alpha * acc[g][i + 3];
					if (post) { o.x = act(o.x); o.y = act(o.y); o.z = act(o.z); o.w = act(o.w); }
					if (vec && i + 3 < valid) *reinterpret_cast<float4 *>(dst + i) = o;
					else {
						if (i + 0 < valid) dst[i + 0] = o.x;
						if (i + 1 < valid) dst[i + 1] = o.y;
						if (i + 2 < valid) dst[i + 2] = o.z;
						if (i + 3 < valid) dst[i + 3] = o.w;
					}
				}
			}
		} else
#pragma unroll
		for (int g = 0; g < NG; g++) {
			const long long col0 = (long long)tn * BN + group_col<CG, TS>(h, g);
			if (P.vecC && col0 + 31 < P.N) {
#pragma unroll
				for (int i = 0; i < 32; i += 4) {
					float4 *cp = reinterpret_cast<float4 *>(crow + col0 + i);
					float4 o;
					if (beta != 0.f) {
						const float4 cv = *cp;
						o.x = fmaf(alpha, acc[g][i + 0], beta * cv.x); o.y = fmaf(alpha, acc[g][i + 1], beta * cv.y);
						o.z = fmaf(alpha, acc[g][i + 2], beta * cv.z); o.w = fmaf(alpha, acc[g][i + 3], beta * cv.w);
					} else {
						o.x = alpha * acc[g][i + 0]; o.y = alpha * acc[g][i + 1];
						o.z = alpha * acc[g][i + 2]; o.w = alpha * acc[g][i + 3];
					}
					if (post) { o.x = act(o.x); o.y = act(o.y); o.z = act(o.z); o.w = act(o.w); }
					*cp = o;
				}
			} else {
#pragma unroll
				for (int i = 0; i < 32; i++) {
					if (col0 + i < P.N) {
						float o = alpha * acc[g][i];
						if (beta != 0.f) o = fmaf(alpha, acc[g][i], beta * crow[col0 + i]);
						crow[col0 + i] = post ? act(o) : o;
					}
				}
			}
		}
	}
#pragma unroll
	for (int g = 0; g < NG; g++) after(g);
}

// PROF compiles the per-role cycle counters in (UGEMM_K1_FLAGS bit 5); the production instantiation has none, which
// keeps ~10 registers out of the epilogue's hot drain loop.
// CONV: the B operand is an image gathered by 4-D TMA boxes (implicit im2col; strides 1..8 through the TMA element stride).  GEMM column n' = io * cv_wp + jo
// with cv_wp = output width rounded up to 32, so every 32-column chunk of a tile is one output-row segment (io, jo0..jo0+31)
// and, for k-block kb = (ki*k + kj) * cv_cblocks + cb, one box {32 channels, 32 x, 1 y, 1 image} of the channels-last copy of
// the image at c = 32*cb, x = jo0*stride + kj - pad, y = io*stride + ki - pad: 32 rows of 128 contiguous bytes, i.e. a quarter of a dense
// K-major B tile.  (TMA needs the box start 16-byte aligned in the contiguous dimension, so the one-pixel shifts of a
// convolution cannot be taken along x of the planar image [measured: illegal instruction]; channels-last puts them on outer
// dimensions.)  Padding pixels and channels beyond ich are TMA out-of-bounds zero fill; columns jo >= wo are computed and
// never stored.
template <int CG, bool PROF, bool CONV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1_3xtf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const K1Params P)
{
	constexpr int BN = 128 * CG;          // accumulator columns (UMMA N)
	constexpr int UMMA_M = 128 * CG;
	constexpr int NG = BN / 2 / 32;       // 32-column groups per epilogue thread
	constexpr uint32_t TMEM_COLS = 2 * BN;

	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
	auto full_bar  = [&](int s) { return bar_base + 8u * s; };
	auto xf_bar    = [&](int s) { return bar_base + 8u * (STAGES + s); };
	auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
	auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
	auto tempty_bar= [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
	const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 4);
	volatile uint32_t *tmem_slot_ptr =
	    reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
	const int cluster_id = (CG == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
	const int num_clusters = (CG == 2) ? (int)num_clusters_x() : (int)gridDim.x;
	const int nkb = P.num_k_blocks;
	const int kc = P.kc_blocks;
	long long *prof = (PROF && P.prof && blockIdx.x < 4) ? P.prof + 16 * blockIdx.x : nullptr;
	const bool heavy = (P.flags & 16384) != 0;

	// ---- one-time setup --------------------------------------------------------------------------------
	if (warp == 0 && lane == 0) {
		prefetch_tmap(&tmA);
		prefetch_tmap(&tmB);
		if (P.tma_store) prefetch_tmap(&tmC);
		for (int s = 0; s < STAGES; s++) {
			mbar_init(full_bar(s), 1);
			mbar_init(xf_bar(s), (XF_SPLIT_STAGE ? 4 * XF_GROUPS : 4) * CG);   // transform warps that publish one stage
			mbar_init(empty_bar(s), 1);
		}
		for (int a = 0; a < 2; a++) {
			mbar_init(tfull_bar(a), 1);
			mbar_init(tempty_bar(a), 8 * CG);  // 8 epilogue warps per CTA of the pair
		}
		for (int d = 0; d < SCHED_SLOTS; d++) {
			mbar_init(bar_base + 8u * (14 + d), 1);
			// consumers of a tile index: TMA thread, 8 transform warps, 8 epilogue warps per CTA + the MMA thread
			mbar_init(bar_base + 8u * (14 + SCHED_SLOTS + d), (1 + 4 * XF_GROUPS + 8) * CG + 1);
		}
		fence_mbar_init();
	}
	__syncwarp();
	if (warp == 1) {
		tmem_alloc<CG>(tmem_slot, TMEM_COLS);
		tmem_relinquish<CG>();
	}
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;

	if (warp < 4) {
		reg_dec<48>();
		if (warp == 0 && lane == 0) {
			// ================= TMA producer =================
			int it = 0;
			long long w_empty = 0; const long long t_begin = tick<PROF>();
			const uint64_t hintA = (P.flags & 128) ? L2_EVICT_LAST : (P.flags & 1024) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
			const uint64_t hintB = (P.flags & 512) ? L2_EVICT_LAST : (P.flags & 256) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
			int nt = 0;
			for (int item; (item = next_tile<CG>(bar_base, nt, false, 0, P.diag, cta_rank, heavy)) >= 0;) {
				for (int sg = 0; sg < 2; sg++) {
				const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
				if (wi.kb1 <= wi.kb0) continue;
				const int tile = wi.tile;
				int tm, tn;
				const int inst = tile / P.tiles_per_batch;
				decode_tile(tile - inst * P.tiles_per_batch, P.tiles_m, P.tiles_n, tm, tn);
				const int a_row0 = tm * UMMA_M + (int)cta_rank * ROWS;
				const int b_row0 = tn * BN + (int)cta_rank * ROWS;
				int cio[ROWS / 32], cjo[ROWS / 32];     // CONV: output row / first output column of each 32-column chunk
				if (CONV) {
#pragma unroll
					for (int j = 0; j < ROWS / 32; j++) {
						const int n0 = b_row0 + 32 * j;
						cio[j] = n0 / P.cv_wp;
						cjo[j] = n0 - cio[j] * P.cv_wp;
					}
				}
				for (int kb = wi.kb0; kb < wi.kb1 && !(P.flags & 64); kb++, it++) {
					const int s = it % STAGES;
					const uint32_t ph = (it / STAGES) & 1;
					const long long tw = tick<PROF>();
					mbar_wait(empty_bar(s), ph ^ 1u, P.diag, 1);
					w_empty += tick<PROF>() - tw;
					mbar_arrive_expect_tx(full_bar(s), RAW_BYTES);
					const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + OPER_BYTES;
					const int k0 = kb * BK;
					if (CONV) {
						const int kpos = kb / P.cv_cblocks, c0 = (kb - kpos * P.cv_cblocks) * 32;
						const int ki = kpos / P.cv_k, kj = kpos - ki * P.cv_k;
						tma_load_3d_hint(sA, &tmA, full_bar(s), k0, a_row0, 0, hintA);       // repacked weights, K-major, shared by all images
#pragma unroll
						for (int j = 0; j < ROWS / 32; j++)
							tma_load_4d_hint(sB + j * 4096, &tmB, full_bar(s), c0, cjo[j] * P.cv_stride + kj - P.cv_pad, cio[j] * P.cv_stride + ki - P.cv_pad, inst, hintB);
						continue;
					}
					if (P.a_kmajor) tma_load_3d_hint(sA, &tmA, full_bar(s), k0, a_row0, inst, hintA);
					else
						for (int j = 0; j < ROWS / 32; j++) tma_load_3d_hint(sA + j * 4096, &tmA, full_bar(s), a_row0 + 32 * j, k0, inst, hintA);
					if (P.b_kmajor) tma_load_3d_hint(sB, &tmB, full_bar(s), k0, b_row0, inst, hintB);
					else
						for (int j = 0; j < ROWS / 32; j++) tma_load_3d_hint(sB + j * 4096, &tmB, full_bar(s), b_row0 + 32 * j, k0, inst, hintB);
				}
				}
			}
			if (prof) { prof[0] = w_empty; prof[1] = tick<PROF>() - t_begin; }
		} else if (warp == 1 && lane == 0 && cta_rank == 0) {
			// ================= MMA issuer (leader CTA) =================
			const uint32_t idesc = idesc_tf32(UMMA_M, BN, P.a_kmajor ? 0 : 1, P.b_kmajor ? 0 : 1);
			// K-major SW128: LBO(enc)=1, SBO=1024 B, k-step (8 fp32) = +32 B inside the swizzle line.
			// MN-major SW128/32B-atom: LBO=4096 B between 32-wide mn groups, SBO=512 B between 4-row
			// k groups, k-step (8 rows) = +1024 B.
			const uint32_t a_lbo = P.a_kmajor ? 1u : 256u, a_sbo = P.a_kmajor ? 64u : 32u, a_lay = P.a_kmajor ? 2u : 1u;
			const uint32_t b_lbo = P.b_kmajor ? 1u : 256u, b_sbo = P.b_kmajor ? 64u : 32u, b_lay = P.b_kmajor ? 2u : 1u;
			const uint32_t a_kstep = P.a_kmajor ? 32u : 1024u, b_kstep = P.b_kmajor ? 32u : 1024u;
			int it = 0, ci = 0;
			long long w_xf = 0, w_te = 0; const long long t_begin = tick<PROF>();
			int nt = 0;
			for (int item; (item = next_tile<CG>(bar_base, nt, false, 0, P.diag, cta_rank, heavy)) >= 0;) {
				for (int sg = 0; sg < 2; sg++) {
				const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
				if (wi.kb1 <= wi.kb0) continue;
				for (int kb0 = wi.kb0; kb0 < wi.kb1; kb0 += kc, ci++) {
					const int acc = ci & 1;
					const uint32_t aph = (ci >> 1) & 1;
					long long tw = tick<PROF>();
					if (CG == 2) mbar_wait_cluster(tempty_bar(acc), aph ^ 1u, P.diag, 2);
					else mbar_wait(tempty_bar(acc), aph ^ 1u, P.diag, 2);
					w_te += tick<PROF>() - tw;
					tc_fence_after();
					const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
					const int kb1 = min(kb0 + kc, wi.kb1);
					for (int kb = kb0; kb < kb1; kb++, it++) {
						const int s = it % STAGES;
						const uint32_t ph = (it / STAGES) & 1;
						tw = tick<PROF>();
						if (!(P.flags & 64)) {
							if (CG == 2) mbar_wait_cluster(xf_bar(s), ph, P.diag, 3);
							else mbar_wait(xf_bar(s), ph, P.diag, 3);
						}
						w_xf += tick<PROF>() - tw;
						tc_fence_after();
						const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + OPER_BYTES;
						const uint32_t sAs = sA + RAW_BYTES, sBs = sB + RAW_BYTES;
#pragma unroll
						for (int k4 = 0; k4 < BK / 8; k4++) {
							const uint64_t dAb = smem_desc(sA + k4 * a_kstep, a_lbo, a_sbo, a_lay);
							const uint64_t dAs = smem_desc(sAs + k4 * a_kstep, a_lbo, a_sbo, a_lay);
							const uint64_t dBb = smem_desc(sB + k4 * b_kstep, b_lbo, b_sbo, b_lay);
							const uint64_t dBs = smem_desc(sBs + k4 * b_kstep, b_lbo, b_sbo, b_lay);
							const uint32_t first = (kb > kb0 || k4 > 0) ? 1u : 0u;
							if (P.flags & 8) { mma_tf32_ss<CG>(d_tmem, dAb, dBb, idesc, first); continue; }
							mma_tf32_ss<CG>(d_tmem, dAs, dBb, idesc, first);
							if (P.flags & 1) {
								mma_tf32_ss_coll<CG, 1>(d_tmem, dAb, dBs, idesc, 1u);
								mma_tf32_ss_coll<CG, 2>(d_tmem, dAb, dBb, idesc, 1u);
							} else {
								mma_tf32_ss<CG>(d_tmem, dAb, dBs, idesc, 1u);
								mma_tf32_ss<CG>(d_tmem, dAb, dBb, idesc, 1u);
							}
						}
						if (!(P.flags & 64)) mma_commit<CG>(empty_bar(s));   // stage free once these MMAs have read it
					}
					mma_commit<CG>(tfull_bar(acc));     // accumulator chunk complete
				}
				}
			}
			if (prof) { prof[2] = w_xf; prof[3] = w_te; prof[4] = tick<PROF>() - t_begin; }
		} else if (warp == 2 && lane == 0 && cta_rank == 0) {
			// ================= tile scheduler (leader CTA) =================
			const uint32_t slots = bar_base + 8u * (14 + 2 * SCHED_SLOTS);
			for (int n = 0;; n++) {
				const int slot = n & (SCHED_SLOTS - 1);
				const uint32_t ph = (n / SCHED_SLOTS) & 1;
				const uint32_t full = bar_base + 8u * (14 + slot), empty = bar_base + 8u * (14 + SCHED_SLOTS + slot);
				if (CG == 2) mbar_wait_cluster(empty, ph ^ 1u, P.diag, 7); else mbar_wait(empty, ph ^ 1u, P.diag, 7);
				int tile;
				if (P.sk_q > 0) {
					// stream-K launches are scheduled statically: full tiles round-robin (sk_full is a multiple of the cluster
					// count, so every pair gets the same number, and tiles of one round are consecutive = L2-friendly), then this
					// pair's own tail range.  The dynamic counter claims up to SCHED_SLOTS items ahead, which at a few tiles per
					// pair would hand the cheap tail ranges to whoever asks last and leave the others with whole tiles.
					const int rounds = P.sk_full / num_clusters;
					tile = n < rounds ? n * num_clusters + cluster_id : (n == rounds && P.sk_full + cluster_id < P.num_tiles ? P.sk_full + cluster_id : -1);
				} else {
					tile = (int)(atomicAdd(P.sched, 1u) - P.sched_base);
					if (tile >= P.num_tiles) tile = -1;
				}
				asm volatile("st.shared.b32 [%0], %1;" ::"r"(slots + 4u * slot), "r"(tile) : "memory");
				if (CG == 2) {
					asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 1;\n\t"
					             "st.shared::cluster.b32 [ra], %1;\n\t}" ::"r"(slots + 4u * slot), "r"(tile) : "memory");
					mbar_arrive_cluster(full, 0);
					mbar_arrive_cluster(full, 1);
				} else {
					mbar_arrive(full);
				}
				if (tile < 0) break;
			}
		}
		__syncwarp();   // reconverge before the .aligned teardown barrier
	} else if (warp < 4 + 4 * XF_GROUPS) {
		// ================= transform warps: write the "small" operand copies =================
		reg_dec<56>();
		const int grp = (warp - 4) >> 2;                 // this warpgroup takes k-blocks with it % XF_GROUPS == grp
		const int t = (threadIdx.x - 128) & 127;
		int it = 0;
		long long w_full = 0, t_work = 0, t_fence = 0; const long long t_begin = tick<PROF>();
		int nt = 0;
		for (int item; (item = next_tile<CG>(bar_base, nt, true, lane, P.diag, cta_rank, heavy)) >= 0;) {
			for (int sg = 0; sg < 2; sg++) {
			const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
			if (wi.kb1 <= wi.kb0) continue;
			for (int kb = wi.kb0; kb < wi.kb1 && !(P.flags & 64); kb++, it++) {
				if (!XF_SPLIT_STAGE && it % XF_GROUPS != grp) continue;
				const int s = it % STAGES;
				const uint32_t ph = (it / STAGES) & 1;
				const long long t0 = tick<PROF>();
				mbar_wait(full_bar(s), ph, P.diag, 4);
				const long long t1 = tick<PROF>();
				const uint32_t raw = smem_base + s * STAGE_BYTES;
#pragma unroll
				for (int half = (XF_SPLIT_STAGE ? grp : 0); half < (XF_SPLIT_STAGE ? grp + 1 : 2); half++) {
					if (P.flags & 4) break;
					float4 v[8];
#pragma unroll
					for (int i = 0; i < 8; i++) v[i] = lds128(raw + (uint32_t)(t + 128 * (half * 8 + i)) * 16u);
					// x = +-Inf: Inf - Inf would make `small` NaN and turn the reference's +-Inf results into NaN; its small part is 0.
					// One test per thread and stage instead of a compare + select per element: the OR of all 32 bit patterns has an
					// all-ones exponent whenever one of them has (false positives only cost the guarded path, which is exact too).
					uint32_t ored = 0;
#pragma unroll
					for (int i = 0; i < 8; i++)
						ored |= __float_as_uint(v[i].x) | __float_as_uint(v[i].y) | __float_as_uint(v[i].z) | __float_as_uint(v[i].w);
					const bool guard = (ored & 0x7F800000u) == 0x7F800000u;
					auto pass = [&](auto guarded) {
#pragma unroll
						for (int i = 0; i < 8; i++) {
							const uint32_t off = (uint32_t)(t + 128 * (half * 8 + i)) * 16u;
							float4 b, sm;
							if (P.split == 0) {
								b.x = tf32_trunc(v[i].x); b.y = tf32_trunc(v[i].y); b.z = tf32_trunc(v[i].z); b.w = tf32_trunc(v[i].w);
							} else {
								b.x = tf32_rna(v[i].x); b.y = tf32_rna(v[i].y); b.z = tf32_rna(v[i].z); b.w = tf32_rna(v[i].w);
							}
							if (decltype(guarded)::value) { sm.x = small_part(v[i].x, b.x); sm.y = small_part(v[i].y, b.y); sm.z = small_part(v[i].z, b.z); sm.w = small_part(v[i].w, b.w); }
							else { sm.x = v[i].x - b.x; sm.y = v[i].y - b.y; sm.z = v[i].z - b.z; sm.w = v[i].w - b.w; }
							if (P.flags & 2) continue;
							sts128(raw + RAW_BYTES + off, sm);
							if (P.split != 0) sts128(raw + off, b);
						}
					};
					if (guard) pass(std::true_type{}); else pass(std::false_type{});
				}
				const long long t2 = tick<PROF>();
				fence_proxy_async_smem();
				__syncwarp();
				if (lane == 0) arrive_on_leader<CG>(xf_bar(s), cta_rank, heavy);
				const long long t3 = tick<PROF>();
				w_full += t1 - t0; t_work += t2 - t1; t_fence += t3 - t2;
			}
			}
		}
		if (prof && threadIdx.x == 128) { prof[5] = w_full; prof[6] = t_work; prof[7] = t_fence; prof[8] = tick<PROF>() - t_begin; }
	} else {
		// ================= epilogue warps =================
		reg_inc<160>();
		const int e = warp - (4 + 4 * XF_GROUPS);
		const int q = e & 3;        // TMEM lane quarter (must equal warp % 4)
		const int h = e >> 2;       // column half
		const float alpha = P.alpha;
		const float bs = P.beta / P.alpha;                 // alpha != 0 here (alpha == 0 never reaches a GEMM kernel)
		const bool preload_c = P.beta != 0.f && fabsf(bs) < 1e18f && fabsf(bs) > 1e-18f;
		int ci = 0;
		long long w_tf = 0, t_drain = 0, t_store = 0; const long long t_begin = tick<PROF>();
		int nt = 0;
		for (int item; (item = next_tile<CG>(bar_base, nt, true, lane, P.diag, cta_rank, heavy)) >= 0;) {
			for (int sg = 0; sg < 2; sg++) {
			const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
			if (wi.kb1 <= wi.kb0) continue;
			const int tile = wi.tile;
			int tm, tn;
			const int inst = tile / P.tiles_per_batch;
			decode_tile(tile - inst * P.tiles_per_batch, P.tiles_m, P.tiles_n, tm, tn);
			float acc[NG][32];
			const long long row = (long long)tm * UMMA_M + (long long)cta_rank * ROWS + q * 32 + lane;
			float *crow = P.C + (long long)inst * P.strideC + row * (CONV ? (long long)P.cv_npix : P.ldc);
			epi_init_acc<CG, CONV, false>(acc, P, wi, preload_c, bs, row, crow, tn, h);
			for (int kb0 = wi.kb0; kb0 < wi.kb1; kb0 += kc, ci++) {
				const int ab = ci & 1;
				const uint32_t aph = (ci >> 1) & 1;
				const long long t0 = tick<PROF>();
				mbar_wait(tfull_bar(ab), aph, P.diag, 5);
				const long long t1 = tick<PROF>();
				tc_fence_after();
				const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + h * (BN / 2));
#pragma unroll
				for (int g = 0; g < 2 * NG; g++) {
					float v[16];
					tmem_ld_32x32b_x16(taddr + g * 16, v);
#pragma unroll
					for (int i = 0; i < 16; i++) acc[g >> 1][(g & 1) * 16 + i] += v[i];   // fp32 round-to-nearest promotion
				}
				tc_fence_before();
				__syncwarp();
				if (lane == 0) arrive_on_leader<CG>(tempty_bar(ab), cta_rank, heavy);
				w_tf += t1 - t0; t_drain += tick<PROF>() - t1;
			}
			const long long ts0 = tick<PROF>();
			epi_store_tile<CG, CONV, false>(acc, P, &tmC, wi, preload_c, alpha, row, crow, tm, tn, inst, q, h, e, lane, cta_rank, bar_base);
			t_store += tick<PROF>() - ts0;
			}
		}
		if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's TMA stores are complete
		if (prof && threadIdx.x == 32 * (4 + 4 * XF_GROUPS)) { prof[9] = w_tf; prof[10] = t_drain; prof[11] = t_store; prof[12] = tick<PROF>() - t_begin; }
	}

	// ---- teardown: everyone (both CTAs of a pair) done before TMEM is returned ------------------------------
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	if (warp == 1) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// ===============================================================================================================
// K1-TS: the same 3xTF32 product with the A operand in TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc, ...).
//
// Why (tools/mma_rate.cu, profiles/r2f_mma_rate3.jsonl): issued back to back, a TF32 MMA whose A operand comes from shared
// memory takes 145 cycles at UMMA 256x256x8 (81-86 at N = 128) -- the rate the SS kernel above runs at -- while the same MMA
// with A in TMEM runs at the instruction floor for every N (128 / 96 / 64 / 32 cycles at N = 256 / 192 / 128 / 64), under
// shared-memory and tcgen05.ld traffic.  So here the transform warps write op(A) -- raw (its TF32 truncation is A_big) and
// small -- straight from registers into TMEM with tcgen05.st; only B keeps a shared-memory "small" copy.
//
// TMEM budget (512 columns).  A stages take 64 columns each (32 raw + 32 small), which leaves no room for two 256-column
// accumulators.  The accumulator is therefore cut into 64-COLUMN SLICES, each accumulated by its own UMMA 256x64x8 (32-cycle
// floor: same tensor throughput as one 256-column MMA), and the promotion schedule of the slices is STAGGERED: with promotion
// every kc = 4 k-blocks, slice j hands its partial sums to the epilogue after k-blocks j, j+4, j+8, ... -- one 64-column slice
// per k-block instead of 256 columns every fourth.  A slice that has been handed over continues in a free buffer, so
// NSL + 1 slice buffers in a FIFO ring (5 x 64 = 320 columns for a 256-wide tile) replace 2 x 256, and 3 A stages fit.
// Both sides count hand-overs with one running index: buffer = index % (NSL + 1).
//
// Shared memory: 4 stages of 48 KiB (A raw | B raw | B small).  With cta_group::2 an N = 64 MMA takes accumulator columns
// 0..31 from the leader's B rows and 32..63 from the peer's, so CTA r loads B in 32-row groups: shared-memory rows 32g..32g+31
// hold columns n0 + 64g + 32r .. +31 of the tile, and accumulator column c of slice g is tile column 64g + c.
// Roles and barriers as in the SS kernel, plus afree[] (TMEM A stage consumed) and per-buffer tfull[] / tempty[].
// ===============================================================================================================
namespace tsk {
constexpr int TS_STAGES = 4;
constexpr int TS_STAGE_BYTES = 3 * OPER_BYTES;         // A raw | B raw | B small = 48 KiB
constexpr int SLICE = 64;                             // accumulator columns per MMA (UMMA N)
constexpr int TS_SMEM_BYTES = TS_STAGES * TS_STAGE_BYTES + 1024 + 8 * CSTAGE_BYTES + 1024;   // ring | barriers | C staging | alignment slack
static_assert(TS_SMEM_BYTES <= 232448, "shared-memory budget of one CTA (227 KiB)");
// barrier block: 8-byte slots counted from bar_base
constexpr int B_FULL = 0, B_XF = 4, B_EMPTY = 8, B_AFREE = 12, B_TFULL = 16, B_TEMPTY = 21;
constexpr int B_SCHED = 26;                           // tile-index ring: full[4], empty[4], 4 x 4-byte slots (next_tile's layout, rebased)
constexpr int B_TMEM = 37;
}

// CONV: the B operand is an image gathered by 4-D TMA boxes (implicit im2col, see the SS kernel): every 32-row group of a B stage
// is one output-row segment, which is exactly the granularity the TS kernel loads B in.
template <int CG, bool PROF, bool CONV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmW, const K1Params P)
{
	using namespace tsk;
	constexpr int BN = 128 * CG, UMMA_M = 128 * CG;
	constexpr int NSL = BN / SLICE, NBUF = NSL + 1;                      // slices per tile, slice buffers in the ring
	constexpr int NA = (512 - NBUF * SLICE) / 64 < TS_STAGES ? (512 - NBUF * SLICE) / 64 : TS_STAGES;   // TMEM A stages: 3 (CG = 2), 4 (CG = 1)
	constexpr uint32_t A_COL0 = NBUF * SLICE;
	constexpr uint32_t SL16 = (SLICE / CG) * 128 / 16;                   // one slice's B rows in this CTA's stage, in 16-byte units
	static_assert(NA >= 2, "need at least two TMEM A stages");

	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + TS_STAGES * TS_STAGE_BYTES;
	auto bar = [&](int idx) { return bar_base + 8u * (uint32_t)idx; };
	const uint32_t sched_bars = bar_base + 8u * (B_SCHED - 14);          // next_tile() addresses its ring at slots 14.. of the base it is given
	const uint32_t tmem_slot = bar(B_TMEM);
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
	const int cluster_id = (CG == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
	const int num_clusters = (CG == 2) ? (int)num_clusters_x() : (int)gridDim.x;
	const int nkb = P.num_k_blocks;
	const int kc = P.kc_blocks;                       // multiple of NSL, or >= nkb (no promotion inside a tile)
	const int step = kc / NSL > 0 ? kc / NSL : 1;     // k-blocks between two hand-overs
	constexpr bool heavy = false;
	long long *prof = (PROF && P.prof && blockIdx.x < 4) ? P.prof + 16 * blockIdx.x : nullptr;   // per-role cycle counters (UGEMM_K1_FLAGS bit 5)

	if (warp == 0 && lane == 0) {
		prefetch_tmap(&tmA);
		prefetch_tmap(&tmB);
		if (P.tma_store) prefetch_tmap(&tmC);
		if (P.sk_q > 0) prefetch_tmap(&tmW);
		for (int s = 0; s < TS_STAGES; s++) {
			mbar_init(bar(B_FULL + s), 1);
			mbar_init(bar(B_XF + s), 8 * CG);        // 8 transform warps per CTA of the pair
			mbar_init(bar(B_EMPTY + s), 1);
			mbar_init(bar(B_AFREE + s), 1);
		}
		for (int b = 0; b < NBUF; b++) {
			mbar_init(bar(B_TFULL + b), 1);
			mbar_init(bar(B_TEMPTY + b), 8 * CG);    // 8 epilogue warps per CTA of the pair
		}
		for (int d = 0; d < SCHED_SLOTS; d++) {
			mbar_init(bar(B_SCHED + d), 1);
			mbar_init(bar(B_SCHED + SCHED_SLOTS + d), (1 + 8 + 8) * CG + 1);   // TMA thread, 8 transform + 8 epilogue warps per CTA, the MMA thread
		}
		fence_mbar_init();
	}
	__syncwarp();
	if (warp == 1) {
		tmem_alloc<CG>(tmem_slot, 512);
		tmem_relinquish<CG>();
	}
	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot_ptr;
	// Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touched nothing an earlier kernel of the
	// stream wrote, so a CTA of this launch may do it while the previous kernel's last tiles are still running on other SMs; from
	// here on the operands, C, the scheduler counter and the stream-K workspace are read, which needs the predecessor finished.
	// The next launch in the stream may be scheduled as soon as this one's CTAs leave their SMs.
	griddep_launch_dependents();
	griddep_wait();

	if (warp < 4) {
		reg_dec<48>();
		if (warp == 0 && lane == 0) {
			// ================= TMA producer =================
			int s = 0; uint32_t ph = 0;
			int nt = 0;
			long long w_empty = 0; const long long t_begin = tick<PROF>();
			for (int item; (item = next_tile<CG>(sched_bars, nt, false, 0, P.diag, cta_rank, heavy)) >= 0;) {
				for (int sg = 0; sg < 2; sg++) {
				const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
				if (wi.kb1 <= wi.kb0) continue;
				int tm, tn;
				const int inst = wi.tile / P.tiles_per_batch;
				decode_tile(wi.tile - inst * P.tiles_per_batch, P.tiles_m, P.tiles_n, tm, tn, P.group);
				const int a_row0 = tm * UMMA_M + (int)cta_rank * ROWS;
				const int b_col0 = tn * BN + (CG == 2 ? 32 * (int)cta_rank : 0);    // group g: + 64 g (pair) / + 32 g (single CTA)
				int cio[ROWS / 32], cjo[ROWS / 32];     // CONV: output row / first output column of each 32-column group
				if (CONV) {
#pragma unroll
					for (int g = 0; g < ROWS / 32; g++) {
						const int n0 = b_col0 + (CG == 2 ? 64 : 32) * g;
						cio[g] = n0 / P.cv_wp;
						cjo[g] = n0 - cio[g] * P.cv_wp;
					}
				}
				// Serpentine K: every other wave of tiles walks its k-blocks downwards.  All pairs of a wave sweep K together, so a wave
				// ends with the high-k blocks of its panels freshest in L2; the next wave shares one operand's panels with it (8 m-tiles
				// share an n sweep) and, walking down, meets them while they are still there (an upward walk finds its first blocks evicted
				// by its own predecessor: the wave's working set is larger than the L2).  Only the load coordinates change -- the other
				// roles count k-blocks -- and the order of a tile's k-blocks is a fixed function of the tile index and the grid size.
				// (Only where panels are re-read from DRAM: a problem that fits the L2, or one whose big operand is streamed once like
				// config 4's, loses with a downward walk -- the L2's 256-byte promotion then fetches the half line already consumed.)
				const bool down = P.serpentine && ((wi.tile / num_clusters) & 1);
				for (int kbi = wi.kb0; kbi < wi.kb1; kbi++) {
					const int kb = down ? wi.kb1 - 1 - (kbi - wi.kb0) : kbi;
					const long long tw = tick<PROF>();
					mbar_wait(bar(B_EMPTY + s), ph ^ 1u, P.diag, 1);
					w_empty += tick<PROF>() - tw;
					mbar_arrive_expect_tx(bar(B_FULL + s), RAW_BYTES);
					const uint32_t sA = smem_base + s * TS_STAGE_BYTES, sB = sA + OPER_BYTES, fb = bar(B_FULL + s);
					const int k0 = kb * BK;
					if (CONV) {
						// k-block kb = (kernel position ki*k + kj, 32-channel block): one box {32 c, 32 x, 1 y, 1 image} per group
						const int kpos = kb / P.cv_cblocks, c0 = (kb - kpos * P.cv_cblocks) * 32;
						const int ki = kpos / P.cv_k, kj = kpos - ki * P.cv_k;
						tma_load_3d_hint(sA, &tmA, fb, k0, a_row0, 0, L2_EVICT_NORMAL);        // repacked weights, K-major, shared by all images
#pragma unroll
						for (int g = 0; g < ROWS / 32; g++)
							tma_load_4d_hint(sB + g * 4096, &tmB, fb, c0, cjo[g] * P.cv_stride + kj - P.cv_pad, cio[g] * P.cv_stride + ki - P.cv_pad, inst, L2_EVICT_NORMAL);
						if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
						continue;
					}
					if (P.a_kmajor) tma_load_3d_hint(sA, &tmA, fb, k0, a_row0, inst, L2_EVICT_NORMAL);
					else
						for (int j = 0; j < ROWS / 32; j++) tma_load_3d_hint(sA + j * 4096, &tmA, fb, a_row0 + 32 * j, k0, inst, L2_EVICT_NORMAL);
#pragma unroll
					for (int g = 0; g < ROWS / 32; g++) {
						const int n = b_col0 + (CG == 2 ? 64 : 32) * g;
						if (P.b_kmajor) tma_load_3d_hint(sB + g * 4096, &tmB, fb, k0, n, inst, L2_EVICT_NORMAL);
						else tma_load_3d_hint(sB + g * 4096, &tmB, fb, n, k0, inst, L2_EVICT_NORMAL);
					}
					if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
				}
				}
			}
			if (prof) { prof[0] = w_empty; prof[1] = tick<PROF>() - t_begin; }
		} else if (warp == 1 && cta_rank == 0) {
			// ================= MMA issuer (leader CTA): the warp stays converged, one elected lane issues =================
			if (elect_one()) {
				const uint32_t idesc = idesc_tf32(UMMA_M, SLICE, 0, P.b_kmajor ? 0 : 1);
				// B descriptors: K-major SW128 (LBO enc 1, SBO 1024 B, k-step +32 B) or MN-major SW128 / 32-byte atom (LBO 4096 B between
				// 32-wide mn groups, SBO 512 B, k-step +1024 B); the high word is constant, the low word carries the start address
				const uint64_t d0 = P.b_kmajor ? smem_desc(0, 1, 64, 2) : smem_desc(0, 256, 32, 1);
				const uint32_t b_hi = (uint32_t)(d0 >> 32), b_lo0 = (uint32_t)d0 + (((smem_base + OPER_BYTES) & 0x3FFFFu) >> 4);
				const uint32_t b_kstep = (P.b_kmajor ? 32u : 1024u) >> 4;
				int s = 0; uint32_t ph = 0;            // shared-memory stage of the next k-block and its phase
				int a = 0;                             // TMEM A stage of the next k-block
				int ab = 0; uint32_t aph = 0;          // next slice buffer of the ring and its phase
				int nt = 0;
				long long w_xf = 0, w_te = 0; const long long t_begin = tick<PROF>();
				for (int item; (item = next_tile<CG>(sched_bars, nt, false, 0, P.diag, cta_rank, heavy)) >= 0;) {
					for (int sg = 0; sg < 2; sg++) {
					const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
					if (wi.kb1 <= wi.kb0) continue;
					const int nseg = wi.kb1 - wi.kb0;
					// slices that hold columns of C at all: the others (a ragged last n-tile, N <= 64 on a 128-wide tile ...) are neither
					// multiplied nor handed over -- their B rows are TMA zero fill and their columns are never stored
					int nact;
					{
						int tm_, tn_;
						const int inst_ = wi.tile / P.tiles_per_batch;
						decode_tile(wi.tile - inst_ * P.tiles_per_batch, P.tiles_m, P.tiles_n, tm_, tn_, P.group);
						nact = (P.N - tn_ * BN + SLICE - 1) / SLICE;
						nact = nact < 1 ? 1 : nact > NSL ? NSL : nact;
					}
					int buf[NSL];                      // ring buffer of each slice
					uint32_t fresh = (1u << NSL) - 1u; // slices whose next MMA starts a new chunk (overwrites its buffer)
					auto take_buffer = [&]() {
						const long long tw = tick<PROF>();
						if (CG == 2) mbar_wait_cluster(bar(B_TEMPTY + ab), aph ^ 1u, P.diag, 2); else mbar_wait(bar(B_TEMPTY + ab), aph ^ 1u, P.diag, 2);
						w_te += tick<PROF>() - tw;
						const int b = ab;
						if (++ab == NBUF) { ab = 0; aph ^= 1u; }
						return b;
					};
					// (the slices take their first buffers one by one inside the first k-block, each just before its first MMA: the previous
					// tile's last hand-overs are still being promoted, and waiting for four free buffers up front would idle the tensor pipe)
#pragma unroll
					for (int j = 0; j < NSL; j++) buf[j] = 0;
					// slice that hands over next, and the k-block after which it does.  The first hand-over of a tile waits kc k-blocks (then one
					// slice every `step`): the epilogue warps are still storing the previous tile, and a full chunk of slack is what the
					// 2 x 256-column scheme gave them.  A slice's first chunk is therefore kc + j * step <= 2 kc - step k-blocks long, all others kc.
					int jo = 0, next_evt = kc - 1;
					for (int t = 0; t < nseg; t++) {
						const long long tw = tick<PROF>();
						if (CG == 2) mbar_wait_cluster(bar(B_XF + s), ph, P.diag, 3); else mbar_wait(bar(B_XF + s), ph, P.diag, 3);
						w_xf += tick<PROF>() - tw;
						tc_fence_after();
						const uint32_t lo_b = b_lo0 + (uint32_t)s * (TS_STAGE_BYTES >> 4);
						const uint32_t a_raw = tmem_base + A_COL0 + (uint32_t)a * 64u, a_small = a_raw + 32u;
#pragma unroll
						for (int j = 0; j < NSL; j++) {
							if (j >= nact) continue;
							if (t == 0) { buf[j] = take_buffer(); tc_fence_after(); }
							const uint32_t d_tmem = tmem_base + (uint32_t)buf[j] * SLICE;
#pragma unroll
							for (int k4 = 0; k4 < BK / 8; k4++) {
								const uint64_t dBb = desc64(lo_b + (uint32_t)j * SL16 + (uint32_t)k4 * b_kstep, b_hi);
								const uint64_t dBs = desc64(lo_b + (uint32_t)j * SL16 + (uint32_t)k4 * b_kstep + (OPER_BYTES >> 4), b_hi);
								mma_tf32_ts<CG>(d_tmem, a_small + 8u * k4, dBb, idesc, (k4 == 0 && ((fresh >> j) & 1u)) ? 0u : 1u);
								mma_tf32_ts<CG>(d_tmem, a_raw + 8u * k4, dBs, idesc, 1u);
								mma_tf32_ts<CG>(d_tmem, a_raw + 8u * k4, dBb, idesc, 1u);
							}
						}
						fresh = 0;
						mma_commit<CG>(bar(B_EMPTY + s));        // shared-memory stage free once these MMAs have read it
						mma_commit<CG>(bar(B_AFREE + a));        // and so is the TMEM A stage
						if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
						if (++a == NA) a = 0;
						if (t == nseg - 1) {
							// end of the tile (or stream-K part): every slice hands over, oldest buffer first
#pragma unroll
							for (int n = 0; n < NSL; n++) {
								if (n >= nact) continue;
								const int j = jo + n < nact ? jo + n : jo + n - nact;
								int b = buf[0];
#pragma unroll
								for (int jj = 1; jj < NSL; jj++) b = (jj == j) ? buf[jj] : b;
								mma_commit<CG>(bar(B_TFULL + b));
							}
						} else if (t == next_evt) {
							// slice jo hands its chunk to the epilogue and continues in the next buffer of the ring
							int b = buf[0];
#pragma unroll
							for (int jj = 1; jj < NSL; jj++) b = (jj == jo) ? buf[jj] : b;
							mma_commit<CG>(bar(B_TFULL + b));
							const int nb = take_buffer();
							tc_fence_after();
#pragma unroll
							for (int jj = 0; jj < NSL; jj++) buf[jj] = (jj == jo) ? nb : buf[jj];
							fresh |= 1u << jo;
							jo = (jo + 1 == nact) ? 0 : jo + 1;
							next_evt += step;
						}
					}
					}
				}
				if (prof) { prof[2] = w_xf; prof[3] = w_te; prof[4] = tick<PROF>() - t_begin; }
			}
			__syncwarp();
		} else if (warp == 2 && lane == 0 && cta_rank == 0) {
			// ================= tile scheduler (leader CTA) =================
			// Item n is claimed once every role has picked up item n-1: a pair never holds more than the tile it works on plus one,
			// so a problem with only a few tiles per pair is shared out evenly (claiming as far ahead as the ring allows let the first
			// pairs to start take four tiles each), while a pair that runs late -- its SMs busy with another kernel -- still claims less.
			// Stream-K launches claim their whole tiles the same way; the pair's own tail range follows when the counter runs dry.
			const uint32_t slots = sched_bars + 8u * (14 + 2 * SCHED_SLOTS);
			const int limit = P.sk_q > 0 ? P.sk_full : P.num_tiles;
			bool tail_given = false;
			for (int n = 0;; n++) {
				const int slot = n & (SCHED_SLOTS - 1);
				const uint32_t full = sched_bars + 8u * (14 + slot);
				if (n >= 1) {
					const uint32_t pempty = sched_bars + 8u * (14 + SCHED_SLOTS + ((n - 1) & (SCHED_SLOTS - 1))), pph = ((n - 1) / SCHED_SLOTS) & 1;
					if (CG == 2) mbar_wait_cluster(pempty, pph, P.diag, 7); else mbar_wait(pempty, pph, P.diag, 7);
				}
				if (n >= SCHED_SLOTS) {
					// the slot's own barrier (item n - 4 read by everyone): implied by the wait above, since roles pick items up in order, and
					// therefore always complete already -- waited on all the same so that the overwrite below is ordered after those reads by
					// the barrier they arrived on, not by transitivity (compute-sanitizer racecheck reports the slot otherwise)
					const uint32_t sempty = sched_bars + 8u * (14 + SCHED_SLOTS + slot), sph = ((n / SCHED_SLOTS) & 1) ^ 1u;
					if (CG == 2) mbar_wait_cluster(sempty, sph, P.diag, 7); else mbar_wait(sempty, sph, P.diag, 7);
				}
				int tile = -1;
				if (!tail_given) {
					tile = (int)(atomicAdd(P.sched, 1u) - P.sched_base);
					if (tile >= limit) {
						tile = (P.sk_q > 0 && P.sk_full + cluster_id < P.num_tiles) ? P.sk_full + cluster_id : -1;
						tail_given = true;
					}
				}
				asm volatile("st.shared.b32 [%0], %1;" ::"r"(slots + 4u * slot), "r"(tile) : "memory");
				if (CG == 2) {
					asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 1;\n\t"
					             "st.shared::cluster.b32 [ra], %1;\n\t}" ::"r"(slots + 4u * slot), "r"(tile) : "memory");
					mbar_arrive_cluster(full, 0);      // release at cluster scope: the peer reads the slot written above
					mbar_arrive_cluster(full, 1);
				} else {
					mbar_arrive(full);
				}
				if (tile < 0) break;
			}
		}
		__syncwarp();   // reconverge before the .aligned teardown barrier
	} else if (warp < 12) {
		// ================= transform warps: op(A) raw + small -> TMEM, B small -> shared memory =================
		reg_dec<56>();
		const int t = (int)threadIdx.x - 128;              // 0..255
		const int grp = t >> 7;                            // warpgroup: k columns [16 grp, 16 grp + 16) of A, half of B
		const int w4 = (t >> 5) & 3;                       // TMEM lane quarter of this warp (= warp % 4)
		const int r = w4 * 32 + lane;                      // row of the A tile this thread moves
		const uint32_t a_tmem = tmem_base + ((uint32_t)(w4 * 32) << 16) + A_COL0 + 16u * (uint32_t)grp;
		int s = 0; uint32_t ph = 0;
		int a = 0; uint32_t aph = 0;
		int nt = 0;
		long long w_full = 0, w_afree = 0, t_fence = 0; const long long t_begin = tick<PROF>();
		for (int item; (item = next_tile<CG>(sched_bars, nt, true, lane, P.diag, cta_rank, heavy)) >= 0;) {
			for (int sg = 0; sg < 2; sg++) {
			const Item wi = decode_item(item, sg, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
			if (wi.kb1 <= wi.kb0) continue;
			for (int kb = wi.kb0; kb < wi.kb1; kb++) {
				const long long t0 = tick<PROF>();
				mbar_wait(bar(B_FULL + s), ph, P.diag, 4);                 // raw tiles have landed
				w_full += tick<PROF>() - t0;
				const uint32_t raw = smem_base + s * TS_STAGE_BYTES;
				float av[16];
				if (P.a_kmajor) {
					// K-major SW128: row r at r * 128, 16-byte chunk c at (c ^ (r & 7)) * 16
#pragma unroll
					for (int c = 0; c < 4; c++) {
						const float4 v = lds128(raw + (uint32_t)r * 128u + (uint32_t)(((4 * grp + c) ^ (r & 7)) << 4));
						av[4 * c + 0] = v.x; av[4 * c + 1] = v.y; av[4 * c + 2] = v.z; av[4 * c + 3] = v.w;
					}
				} else {
					// MN-major SW128 / 32-byte atom: 32-row group w4 at w4 * 4096, k line at k * 128, 32-byte atom (lane / 8) ^ (k & 3)
#pragma unroll
					for (int kk = 0; kk < 16; kk++) {
						const int k = 16 * grp + kk;
						av[kk] = lds32(raw + (uint32_t)w4 * 4096u + (uint32_t)k * 128u + (uint32_t)((((lane >> 3) ^ (k & 3)) << 5) + ((lane & 7) << 2)));
					}
				}
				float4 bv[4];
#pragma unroll
				for (int i = 0; i < 4; i++) bv[i] = lds128(raw + OPER_BYTES + (uint32_t)(t + 256 * i) * 16u);
				// x = +-Inf: Inf - Inf would make `small` NaN; its small part is 0.  One test per operand, thread and stage (see the SS kernel).
				uint32_t ored = 0;
#pragma unroll
				for (int i = 0; i < 16; i++) ored |= __float_as_uint(av[i]);
				const bool guard_a = (ored & 0x7F800000u) == 0x7F800000u;
				const long long t1 = tick<PROF>();
				mbar_wait(bar(B_AFREE + a), aph ^ 1u, P.diag, 8);          // the MMAs that read this TMEM A stage last have retired
				w_afree += tick<PROF>() - t1;
				tc_fence_after();
				tmem_st_32x32b_x16(a_tmem + (uint32_t)a * 64u, av);        // raw: the tensor core truncates it to A_big itself
				if (guard_a) {
#pragma unroll
					for (int i = 0; i < 16; i++) av[i] = small_part(av[i], tf32_trunc(av[i]));
				} else {
#pragma unroll
					for (int i = 0; i < 16; i++) av[i] -= tf32_trunc(av[i]);
				}
				tmem_st_32x32b_x16(a_tmem + (uint32_t)a * 64u + 32u, av);
				ored = 0;
#pragma unroll
				for (int i = 0; i < 4; i++) ored |= __float_as_uint(bv[i].x) | __float_as_uint(bv[i].y) | __float_as_uint(bv[i].z) | __float_as_uint(bv[i].w);
				const bool guard_b = (ored & 0x7F800000u) == 0x7F800000u;
#pragma unroll
				for (int i = 0; i < 4; i++) {
					float4 sm;
					if (guard_b) {
						sm.x = small_part(bv[i].x, tf32_trunc(bv[i].x)); sm.y = small_part(bv[i].y, tf32_trunc(bv[i].y));
						sm.z = small_part(bv[i].z, tf32_trunc(bv[i].z)); sm.w = small_part(bv[i].w, tf32_trunc(bv[i].w));
					} else {
						sm.x = bv[i].x - tf32_trunc(bv[i].x); sm.y = bv[i].y - tf32_trunc(bv[i].y);
						sm.z = bv[i].z - tf32_trunc(bv[i].z); sm.w = bv[i].w - tf32_trunc(bv[i].w);
					}
					sts128(raw + 2 * OPER_BYTES + (uint32_t)(t + 256 * i) * 16u, sm);
				}
				const long long t2 = tick<PROF>();
				tmem_st_wait();
				fence_proxy_async_smem();
				tc_fence_before();
				__syncwarp();
				if (lane == 0) arrive_on_leader<CG>(bar(B_XF + s), cta_rank, heavy);
				t_fence += tick<PROF>() - t2;
				if (++s == TS_STAGES) { s = 0; ph ^= 1u; }
				if (++a == NA) { a = 0; aph ^= 1u; }
			}
			}
		}
		if (prof && threadIdx.x == 128) { prof[5] = w_full; prof[6] = w_afree; prof[7] = t_fence; prof[8] = tick<PROF>() - t_begin; }
	} else {
		// ================= epilogue warps =================
		reg_inc<160>();
		constexpr int NG = NSL;                            // one 32-column group per slice and thread
		const int e = warp - 12;
		const int q = e & 3;        // TMEM lane quarter (must equal warp % 4)
		const int h = e >> 2;       // column half of every slice
		// beta != 0: the old C is folded in UP FRONT, as in the SS kernel, but with the weighting turned round: the running sums of a
		// tile start at C itself (a pure load) and every promotion adds (alpha/beta) * partial sums (one FMA instead of one add), the
		// tile end multiplies by beta.  A stream-K part starts at zero and stays unweighted (the fix-up pass applies alpha and beta).
		const float ab = P.alpha / P.beta;
		const bool preload_c = P.beta != 0.f && fabsf(ab) < 1e18f && fabsf(ab) > 1e-18f;
		int db = 0; uint32_t dph = 0;                      // next slice buffer to be handed over, and its phase
		long long w_tf = 0, t_store = 0; const long long t_begin = tick<PROF>();
		float acc[NG][32];
		// promote slice j: add the 32 columns of this thread's half of the handed-over buffer into the running fp32 sums
		auto drain = [&](int j, float r) {
			const long long tw = tick<PROF>();
			mbar_wait(bar(B_TFULL + db), dph, P.diag, 5);
			w_tf += tick<PROF>() - tw;
			tc_fence_after();
			const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db * SLICE + h * 32);
#pragma unroll
			for (int half = 0; half < 2; half++) {
				float v[16];
				tmem_ld_32x32b_x16(taddr + 16 * half, v);
#pragma unroll
				for (int jj = 0; jj < NSL; jj++)
					if (jj == j) {
#pragma unroll
						for (int i = 0; i < 16; i++) acc[jj][16 * half + i] = fmaf(r, v[i], acc[jj][16 * half + i]);   // fp32 round-to-nearest promotion (r = 1: an add)
					}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) arrive_on_leader<CG>(bar(B_TEMPTY + db), cta_rank, heavy);
			if (++db == NBUF) { db = 0; dph ^= 1u; }
		};
		// The work items of this CTA as a stream of segments (a whole tile, or one part of a stream-K range).  Hand-over number ev of
		// a segment always belongs to slice ev % nact (natural hand-overs go round the active slices, the final ones continue the round).
		struct Seg { Item wi; int tm, tn, inst, nact; };    // (kept small: two of them live beside 128 accumulator registers); nact: see the MMA thread
		int nt = 0, item = -1, sgn = 2;
		auto fetch = [&](Seg &sg) -> bool {
			for (;;) {
				if (sgn >= 2) {
					item = next_tile<CG>(sched_bars, nt, true, lane, P.diag, cta_rank, heavy);
					sgn = 0;
					if (item < 0) return false;
				}
				sg.wi = decode_item(item, sgn++, P.sk_full, P.sk_rem, P.sk_nch, P.sk_q, kc, nkb);
				if (sg.wi.kb1 > sg.wi.kb0) break;
			}
			sg.inst = sg.wi.tile / P.tiles_per_batch;
			decode_tile(sg.wi.tile - sg.inst * P.tiles_per_batch, P.tiles_m, P.tiles_n, sg.tm, sg.tn, P.group);
			sg.nact = (P.N - sg.tn * BN + SLICE - 1) / SLICE;
			sg.nact = sg.nact < 1 ? 1 : sg.nact > NSL ? NSL : sg.nact;
			return true;
		};
		auto row_of = [&](const Seg &sg) { return (long long)sg.tm * UMMA_M + (long long)cta_rank * ROWS + q * 32 + lane; };
		auto crow_of = [&](const Seg &sg) { return P.C + (long long)sg.inst * P.strideC + row_of(sg) * (CONV ? (long long)P.cv_npix : P.ldc); };
		// hand-overs of a segment: after k-blocks kc-1, kc-1+step, ... (not the last one), and one per active slice at the end
		auto nev_of = [&](const Seg &sg) { const int nseg = sg.wi.kb1 - sg.wi.kb0; return (nseg - 1 >= kc ? (nseg - 1 - kc) / step + 1 : 0) + sg.nact; };
		// beta != 0: the old C is folded in up front (running sums start at (beta/alpha) * C), see the SS kernel
		auto weighted = [&](const Seg &sg) { return !CONV && preload_c && sg.wi.slot < 0; };     // this segment's sums are in units of beta
		auto from_c = [&](const Seg &sg) { return weighted(sg) && row_of(sg) < P.M; };
		auto weight = [&](const Seg &sg) { return weighted(sg) ? ab : 1.f; };
		Seg cur, nxt;
		bool have = fetch(cur);
		if (have) {
#pragma unroll
			for (int g = 0; g < NG; g++) epi_init_group<CG, true>(acc[g], g, P, from_c(cur), crow_of(cur), cur.tn, h);
		}
		int done = 0;                 // hand-overs of `cur` taken early, while the previous segment was being stored
		while (have) {
			const int nev = nev_of(cur);
			const float r_cur = weight(cur);
			for (int ev = done; ev < nev; ev++) drain(ev % cur.nact, r_cur);
			const bool have_next = fetch(nxt);
			done = 0;
			if (have_next && from_c(nxt)) {
				// beta != 0: the next segment's old C is needed group by group during the store below; start it towards L2 now, so that
				// those loads are L2 hits instead of four DRAM round trips in a row on the path that gives slice buffers back
				const float *c0 = crow_of(nxt) + (long long)nxt.tn * BN;
#pragma unroll
				for (int g = 0; g < NG; g++) {
					const long long col0 = (long long)nxt.tn * BN + group_col<CG, true>(h, g);
					if (col0 < P.N) { prefetch_l2(c0 + group_col<CG, true>(h, g)); if (col0 + 31 < P.N) prefetch_l2(c0 + group_col<CG, true>(h, g) + 31); }
				}
			}
			// Store `cur` one 32-column group at a time.  As soon as group g has been staged its registers are re-armed for the next
			// segment, and hand-overs of the next segment that are already waiting (slice <= g) are taken at once: the MMA thread
			// needs their buffers back within a few k-blocks, a whole-tile store takes longer than that.
			const float r_nxt = have_next ? weight(nxt) : 1.f;
			const int nev_nxt = have_next ? nev_of(nxt) : 0;
			auto after_group = [&](int g) {
				if (!have_next) return;
#pragma unroll
				for (int gg = 0; gg < NG; gg++)
					if (gg == g) epi_init_group<CG, true>(acc[gg], gg, P, from_c(nxt), crow_of(nxt), nxt.tn, h);
				while (done <= g && done < nev_nxt && mbar_try_wait(bar(B_TFULL + db), dph)) { drain(done % nxt.nact, r_nxt); done++; }
			};
			const long long ts0 = tick<PROF>();
			epi_store_tile<CG, CONV, true>(acc, P, &tmC, cur.wi, preload_c, weighted(cur) ? P.beta : P.alpha, row_of(cur), crow_of(cur), cur.tm, cur.tn, cur.inst, q, h, e, lane, cta_rank, bar_base, after_group, P.sk_q > 0 ? &tmW : nullptr);
			t_store += tick<PROF>() - ts0;
			cur = nxt;
			have = have_next;
		}
		if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's TMA stores are complete
		if (prof && threadIdx.x == 32 * 12) { prof[9] = w_tf; prof[10] = 0; prof[11] = t_store; prof[12] = tick<PROF>() - t_begin; }
	}

	tc_fence_before();
	if (CG == 2) { cluster_arrive(); cluster_wait(); } else __syncthreads();
	if (warp == 1) tmem_dealloc<CG>(tmem_base, 512);
}

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
template <int CG, bool CONV, bool TS = false>
cudaError_t launch_with_tail(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, K1Params &P, long long nt, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
	const int kc_eff = (t.kc_blocks > 0 && t.kc_blocks < P.num_k_blocks) ? t.kc_blocks : P.num_k_blocks;
	const int nch = (P.num_k_blocks + kc_eff - 1) / kc_eff;
	const long long pairs = sm_count / CG;
	const int tile_m = 128 * CG, tile_n = 128 * CG;
	long long items = nt;
	float *ws = nullptr;
	if (!(t.flags & 2048) && nch >= 2 && pairs > 0 && nt % pairs != 0) {
		const long long full = nt / pairs * pairs, rem = nt - full;
		const long long q = (rem * nch + pairs - 1) / pairs, ranges = (rem * nch + q - 1) / q;
		// expected saving: (1 - q/nch) of one round out of ceil(nt / pairs); the fix-up pass and the tail's poorer L2 locality
		// (parts of one tile run at different k offsets) cost a few percent of a round, so small savings are not worth it
		const double saved_rounds = 1.0 - (double)q / nch, rounds = (double)((nt + pairs - 1) / pairs);
		// (measured: a modelled saving of 7-11 % of the launch came out as a 2-4 % loss, 13 % as a 10 % gain)
		// TS kernel [measured, profiles/r2n_sk_sweep.jsonl, r2o_sk_ablate.txt]: once at least one full round precedes the tail, the part
		// stores, the fix-up pass and the tail's colder start cost about as much as 30 k-blocks of a pair; below that the tail is a loss
		// (4096 x 3072 x 2048: 24 k-blocks saved, 219 vs 209 us), above it a gain (2560^3: 48 saved, 155 vs 175 us; 4096^3: 68, 494 vs 522)
		// ... and at least 2 % of a pair's whole work: at 8192^3 the tail saves 40 of 3543 k-blocks per pair, costs 0.36 GB of extra DRAM
		// traffic (parts, colder tail) and measures as nothing (3.66 vs 3.64 ms, profiles/r3a_traffic.csv)
		const long long saved_kb = (nch - q) * (long long)kc_eff, pair_kb = (nt * (long long)P.num_k_blocks + pairs - 1) / pairs;
		const bool worth = TS ? (full == 0 ? saved_rounds >= 0.15 : (saved_kb >= 32 && saved_kb * 50 >= pair_kb))
		                      : (saved_rounds >= 0.15 && saved_rounds / rounds >= 0.12);
		if ((worth || ((t.flags & 131072) && saved_rounds > 0.0)) && rem * nch < 0x3fffffffLL) {      // (bit 17: tail whenever it saves anything, A/B runs)
			const size_t tile_bytes = (size_t)tile_m * tile_n * sizeof(float);
			if (cudaMallocAsync(reinterpret_cast<void **>(&ws), (size_t)(2 * ranges) * tile_bytes, stream) == cudaSuccess) {
				P.sk_full = (int)full; P.sk_rem = (int)rem; P.sk_nch = nch; P.sk_q = (int)q; P.sk_ws = ws;
				items = full + ranges;
			} else { cudaGetLastError(); ws = nullptr; }
		}
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
	P.vecC = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 4 == 0) ? 1 : 0;
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

cudaError_t launch_k1_3xtf32(const Problem &p, const K1Tuning &t, cudaStream_t stream, int sm_count)
{
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
	if (cg == 1) return launch_cg<1>(p, t, stream, sm_count);
	return launch_cg<2>(p, t, stream, sm_count);
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
