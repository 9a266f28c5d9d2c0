// k1_common.cuh -- what the two implementations of K1 (k1_ss.cuh: both operands from shared memory, round 1; k1_ts.cuh: op(A) in
// tensor memory, the production kernel) and their host side (k1_tcgen05.cu) share: stage geometry, the kernel parameter block,
// watchdogged mbarrier waits, tile order and stream-K work items, the dynamic tile-index ring, the 3xTF32 split, and the
// epilogue pieces (running-sum initialisation, fused alpha/beta/bias/LeakyReLU + TMA / strided stores, stream-K part stores).
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <type_traits>

namespace ugemm {
namespace {      // internal linkage: these headers are included by exactly one translation unit, k1_tcgen05.cu

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
		if (P.vecC && col0 + 31 < P.N && !(P.flags & 1048576)) {
			// One row per thread: a warp's load instruction touches 32 different lines whatever its width, and the L1 takes one
			// wavefront per line -- so 256-bit loads (sm_100: LDG.E.256) halve the wavefronts of the preload against 128-bit ones
			// (4 x 32 instead of 8 x 32 per warp and group).  A row that starts on an odd 16-byte boundary (ldc % 8 == 4) takes
			// 16 + 3 x 32 + 16 bytes; the two kinds of rows alternate within a warp, each load then touches 16 lines.
			const float *p = crow + col0;
			if ((reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
#pragma unroll
				for (int i = 0; i < 32; i += 8) ldg256(p + i, &a[i]);
			} else {
				const float4 c0 = *reinterpret_cast<const float4 *>(p);
				a[0] = c0.x; a[1] = c0.y; a[2] = c0.z; a[3] = c0.w;
#pragma unroll
				for (int i = 4; i < 28; i += 8) ldg256(p + i, &a[i]);
				const float4 c7 = *reinterpret_cast<const float4 *>(p + 28);
				a[28] = c7.x; a[29] = c7.y; a[30] = c7.z; a[31] = c7.w;
			}
		} else if (P.vecC && col0 + 31 < P.N) {
			// (flags bit 20, A/B runs: the 128-bit form)
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

} // namespace
} // namespace ugemm
