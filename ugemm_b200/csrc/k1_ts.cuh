// k1_ts.cuh -- K1, production implementation ("TS"): op(A) lives in tensor memory.
#pragma once
#include "k1_common.cuh"

namespace ugemm {
namespace {      // internal linkage: these headers are included by exactly one translation unit, k1_tcgen05.cu

using namespace ptx;

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
// NSL + 1 slice buffers in a FIFO ring (5 x 64 = 320 columns for a 256-wide tile) would replace 2 x 256 and leave room for 3 A
// stages; the pair kernel runs NSL + 2 buffers and 2 A stages (see UGEMM_TS_XBUF2 below for the measurement).
// Both sides count hand-overs with one running index: buffer = index % NBUF.
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
constexpr int B_FULL = 0, B_XF = 4, B_EMPTY = 8, B_AFREE = 12, B_TFULL = 16, B_TEMPTY = 22;      // (room for 6 slice buffers)
constexpr int B_SCHED = 28;                           // tile-index ring: full[4], empty[4], 4 x 4-byte slots (next_tile's layout, rebased)
constexpr int B_TMEM = 39;
// Slice buffers beyond NSL + 1 (build-time knobs; single CTA / CTA pair).  A pair runs SIX buffers and TWO TMEM A stages (6 x 64 +
// 2 x 64 = 512 columns) instead of five and three: at a tile boundary the four final hand-overs leave two buffers instead of one for
// the next tile's first slices, and in the steady state the MMA thread -- which issues about a k-block ahead of the tensor pipe --
// finds a free buffer without waiting for the drain of the hand-over before.  [measured] one box, alternating processes
// (profiles/r4j_xbuf_ab.jsonl, r4k_xbuf_ab2.jsonl): c3 NT 0.2106 -> 0.2064 ms (beta 0.5: 0.2222 -> 0.2172), TN / TT -1.5 %, c4 -0.6 %
// (beta = 1 -0.8 %), 16384 x 8192 x 2048 beta = 1 -0.9 %, fused convolution of config 4's layer -1 %, 2048^3 -0.7 %; 4096^3 NN and
// 8192^3 unchanged; the one loser is a long-K product with an MN-major A, whose slower transform misses the third A stage
// (4096^3 TN +1.2 %).  A single CTA (448 of 512 columns in use) gains nothing from a fourth buffer (1024^3, 1536^3, 2560^3 forced to
// single CTAs, 200704 x 128 x 1152: all within 0.5 %), so it keeps three.
#ifndef UGEMM_TS_XBUF1
#define UGEMM_TS_XBUF1 0
#endif
#ifndef UGEMM_TS_XBUF2
#define UGEMM_TS_XBUF2 1
#endif
constexpr int B_CLOAD = 40;                           // one per epilogue warp: the old C of a 32 x 32 box has landed in the warp's staging box (beta != 0)
static_assert(B_XF >= B_FULL + TS_STAGES && B_EMPTY >= B_XF + TS_STAGES && B_AFREE >= B_EMPTY + TS_STAGES && B_TFULL >= B_AFREE + TS_STAGES &&
              B_TEMPTY >= B_TFULL + 6 && B_SCHED >= B_TEMPTY + 6 && B_TMEM >= B_SCHED + 2 * SCHED_SLOTS + 2 && B_CLOAD > B_TMEM &&
              8 * (B_CLOAD + 8) <= 1024, "barrier block: slots overlap or leave the 1 KiB reserved for them");
}

// CONV: the B operand is an image gathered by 4-D TMA boxes (implicit im2col, see the SS kernel): every 32-row group of a B stage
// is one output-row segment, which is exactly the granularity the TS kernel loads B in.
template <int CG, bool PROF, bool CONV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k1ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmW, const K1Params P)
{
	using namespace tsk;
	constexpr int BN = 128 * CG, UMMA_M = 128 * CG;
	constexpr int NSL = BN / SLICE, NBUF = NSL + 1 + (CG == 1 ? UGEMM_TS_XBUF1 : UGEMM_TS_XBUF2);   // slices per tile, slice buffers in the ring
	static_assert(NBUF <= 6, "barrier block has room for six slice buffers");
	constexpr int NA = (512 - NBUF * SLICE) / 64 < TS_STAGES ? (512 - NBUF * SLICE) / 64 : TS_STAGES;   // TMEM A stages: 2 (CG = 2, six slice buffers), 4 (CG = 1)
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
		for (int w = 0; w < 8; w++) mbar_init(bar(B_CLOAD + w), 1);
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
		// Where the old C comes from.  With a TMA-addressable C it is fetched by the TMA unit, one 32 x 32 box at a time, into the warp's
		// staging box (idle between two tile stores) and ADDED to the running sums between two promotions, any time before the tile
		// ends (the sums are in units of beta either way).  Loading it into the registers with global loads when a group is re-armed
		// (flags bit 22, and any C the TMA unit cannot address) costs the load/store pipe 32 wavefronts per warp instruction -- one row
		// per thread -- which the transform warps' shared-memory traffic has to share: 10 % on config 3.
		const bool c_tma = preload_c && !CONV && P.tma_store && !(P.flags & 4194304);
		const uint32_t cbar = bar(B_CLOAD + e), cbox = bar_base + 1024u + (uint32_t)e * CSTAGE_BYTES;
		uint32_t cph = 0;
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
		auto from_c = [&](const Seg &sg) { return weighted(sg) && !c_tma && row_of(sg) < P.M; };             // registers start at C (global loads)
		auto box_row0 = [&](const Seg &sg) { return sg.tm * UMMA_M + (int)cta_rank * ROWS + q * 32; };
		// c_tma: how many of this warp's 32-column groups hold elements of C (group g starts at tile column 64 g + 32 h; warp-uniform)
		auto c_groups = [&](const Seg &sg) {
			if (!c_tma || !weighted(sg) || box_row0(sg) >= P.M) return 0;
			const int rem = P.N - sg.tn * BN - h * 32;
			return rem <= 0 ? 0 : (rem + 63) / 64 > NG ? NG : (rem + 63) / 64;
		};
		auto c_issue = [&](const Seg &sg, int g) {
			if (lane == 0) {
				bulk_wait_group_read0();                       // the box's last TMA store has left shared memory
				mbar_arrive_expect_tx(cbar, CSTAGE_BYTES);     // (a box that sticks out of C is zero-filled and counts in full)
				tma_load_3d_hint(cbox, &tmC, cbar, sg.tn * BN + group_col<CG, true>(h, g), box_row0(sg), sg.inst, L2_EVICT_NORMAL);
			}
		};
		auto c_take = [&](int g) {
			mbar_wait(cbar, cph, P.diag, 9);
			cph ^= 1u;
#pragma unroll
			for (int i = 0; i < 32; i += 4) {
				const float4 v = lds128(cbox + (uint32_t)lane * 128u + (uint32_t)(((i >> 2) ^ (lane & 7)) << 4));
#pragma unroll
				for (int gg = 0; gg < NG; gg++)
					if (gg == g) { acc[gg][i + 0] += v.x; acc[gg][i + 1] += v.y; acc[gg][i + 2] += v.z; acc[gg][i + 3] += v.w; }
			}
			__syncwarp();                                      // every lane has read the box before lane 0 lets the TMA unit overwrite it
		};
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
			// c_tma: after a promotion, take the box asked for after the previous one and ask for the next.  Reproducibility: the place
			// of "+ C" in a group's chain of roundings must not depend on timing, so it is tied to the hand-over NUMBER of the tile --
			// group g is asked for after hand-over NG + g and added after hand-over NG + g + 1 (or after the last one of a short tile);
			// up to NG hand-overs may have been taken early during the previous store (`done`), never more.  (Later places in the tile
			// measure the same: hand-over 12 or 24 instead of NG, profiles/r4e_beta_ab.jsonl.)
			const int ncg = c_groups(cur);
			int cg = 0;
			bool cpend = false;
			auto c_step = [&]() {
				if (cpend) { c_take(cg); cg++; cpend = false; }
				if (cg < ncg) { c_issue(cur, cg); cpend = true; }
			};
			for (int ev = done; ev < nev; ev++) {
				drain(ev % cur.nact, r_cur);
				if (ev >= NG && cg < ncg) c_step();
			}
			while (cg < ncg) c_step();
			const bool have_next = fetch(nxt);
			done = 0;
			if (have_next && from_c(nxt)) {
				// register path of beta != 0: the next segment's old C is needed group by group during the store below; start it towards
				// L2 now, so that those loads are L2 hits instead of four DRAM round trips in a row on the path that gives slice buffers
				// back.  (The TMA path asks a k-block ahead and gains nothing from a prefetch: 0.2217 vs 0.2235 ms on config 3 with one
				// cp.async.bulk.prefetch.tensor per box, profiles/r4e_beta_ab.jsonl.)
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

} // namespace
} // namespace ugemm
