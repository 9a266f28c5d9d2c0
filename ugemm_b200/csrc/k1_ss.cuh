// k1_ss.cuh -- K1, round-1 implementation ("SS"): A and B operands of every tcgen05.mma come from shared memory, two 256-column
// accumulators in TMEM.  Kept for A/B runs (sgemm_cuda_set_k1_variant(1), UGEMM_K1_FLAGS bit 15) and for the RNA-split experiment;
// the role map is in the header comment of k1_tcgen05.cu, the measurements in DESIGN.md sections 3.2-3.4.
#pragma once
#include "k1_common.cuh"

namespace ugemm {
namespace {      // internal linkage: these headers are included by exactly one translation unit, k1_tcgen05.cu

using namespace ptx;

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

} // namespace
} // namespace ugemm
