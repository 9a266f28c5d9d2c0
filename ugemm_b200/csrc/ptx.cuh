// ptx.cuh -- inline-PTX wrappers for the sm_100a features K1 uses: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), cluster primitives, setmaxnreg.  Hand-written; the bit
// layouts of the shared-memory and instruction descriptors follow the PTX ISA "tcgen05" chapter.
#pragma once
#include <cstdint>

namespace ugemm { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { uint32_t l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t num_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
// arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank)
{
	asm volatile("{\n\t.reg .b32 ra;\n\t"
	             "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
	             "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
	             ::"r"(bar), "r"(rank) : "memory");
}
// Same arrive WITHOUT the cluster-scope release: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive, a
// fence that waits for every outstanding memory operation of the thread (hundreds of cycles; 43 % of a transform warp's time in
// the round-1 kernel, ncu source page).  The default (.release at CTA scope) is enough wherever the data being handed over
// does not travel through the generic proxy to the other CTA: shared-memory stages made visible by fence.proxy.async and read by
// the tensor core of the CTA that wrote them, TMEM buffers ordered by tcgen05.fence, plain "slot consumed" notifications.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank)
{
	asm volatile("{\n\t.reg .b32 ra;\n\t"
	             "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
	             "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
	             ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\t"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	             "selp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok;
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\t"
	             "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
	             "selp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok;
}

// ---- proxies / fences -------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- TMA ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const void *tmap)
{ asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory"); }
// 2-D tiled load global -> this CTA's shared memory; completes `bytes of box` on `bar` (this CTA's)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// same with an L2 eviction-priority hint (createpolicy encodings: evict_first / evict_last)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull, L2_EVICT_NORMAL = 0x1000000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1, uint64_t hint)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "l"(hint) : "memory");
}

// 3-D variant (third coordinate = batch instance)
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1, int c2, uint64_t hint)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint) : "memory");
}

// ---- TMA store (shared -> global), bulk-group completion ------------------------------------------------
// 3-D box store: coordinates {c0 = column, c1 = row, c2 = batch instance}; out-of-bounds parts of the box are clipped,
// so ragged tile edges need no guards.  The source must have been made visible to the async proxy (fence.proxy.async).
__device__ __forceinline__ void tma_store_3d(const void *tmap, uint32_t src, int c0, int c1, int c2)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
	             ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, uint32_t src, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
	             ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// 4-D box store (fused convolution: x, y, output channel, image)
__device__ __forceinline__ void tma_store_4d(const void *tmap, uint32_t src, int c0, int c1, int c2, int c3)
{
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
	             ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed bulk groups have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// 4-D variant (implicit-GEMM convolution: x, y, channel, image)
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t hint)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint) : "memory");
}

// ---- tensor memory --------------------------------------------------------------------------------------
template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
	if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
	else         asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int CG> __device__ __forceinline__ void tmem_relinquish()
{
	if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	else         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
	if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
	else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- tcgen05.mma, kind::tf32, both operands from shared memory ---------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc];  accumulate == 0 overwrites D.
template <int CG> __device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	if (CG == 1)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (K-major only): D[tmem] (+)= A[tmem] * B[smem desc]
template <int CG> __device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	if (CG == 1)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same with the A-operand collector: KEEP = latch this A in the tensor core's collector after reading it from
// shared memory (.collector::a::fill), REUSE = take A from the collector instead of shared memory and release it
// (.collector::a::lastuse).  Lets big*small and big*big of one k-step share a single shared-memory read of A_big.
template <int CG, int COLL /*1 fill, 2 lastuse*/>
__device__ __forceinline__ void mma_tf32_ss_coll(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	if (CG == 1 && COLL == 1)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::1.kind::tf32.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else if (CG == 1)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::1.kind::tf32.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else if (COLL == 1)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::2.kind::tf32.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
		             "tcgen05.mma.cta_group::2.kind::tf32.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
		             ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit all previously issued MMAs of this thread to an mbarrier (arrive::one when they retire).
// CG==2: the arrive is multicast to the barrier at the same offset in both CTAs of the pair.
template <int CG> __device__ __forceinline__ void mma_commit(uint32_t bar)
{
	if (CG == 1)
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
	else
		asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
		             ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// shared-memory matrix descriptor (64 bit): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | base_offset [49,52)=0 | lbo_mode [52]=0 | layout type [61,64): 0 none, 1 128B(32B atom), 2 128B, 4 64B, 6 32B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_enc, uint32_t sbo_enc, uint32_t layout)
{
	return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)lbo_enc << 16) | ((uint64_t)sbo_enc << 32) |
	       (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (32 bit) for kind::tf32 with fp32 accumulation:
// c_format=F32 (1) [4,6) | a_format=TF32 (2) [7,10) | b_format=TF32 (2) [10,13) | a_major [15] | b_major [16]
// (0 = K-major, 1 = MN-major) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major)
{
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
	       ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns, thread t gets lane (base+t) ----------------
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32])
{
	uint32_t r[32];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
	      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
	      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16])
{
	uint32_t r[16];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	    : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
	for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---- registers -> TMEM: thread t writes 16 consecutive 32-bit columns of lane (base+t) ---------------------------------
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const float (&v)[16])
{
	asm volatile(
	    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
	    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
	    ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
	      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
	      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
	      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
	    : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one lane of a converged warp (the compiler then issues tcgen05 instructions under that predicate without an election loop)
__device__ __forceinline__ bool elect_one()
{
	uint32_t p;
	asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
	return p != 0;
}
// 64-bit descriptor from its two words (the low word carries the start address >> 4, so stepping a descriptor is one 32-bit add)
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }

// ---- register re-budgeting between warp roles (whole warpgroup must execute it) -----------------------------
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// bring the line holding `p` into L2 (no register, no scoreboard: the later load finds it there)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 256-bit global load (sm_100: LDG.E.256), 32-byte aligned address; dst = 8 consecutive floats
__device__ __forceinline__ void ldg256(const float *p, float *dst)
{
	asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	             : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]), "=f"(dst[7]) : "l"(p));
}

// ---- programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start before its
// predecessor in the stream has finished; nothing the predecessor wrote may be touched before griddep_wait() returns.  The
// predecessor lets its dependents be scheduled with griddep_launch_dependents() (else: when it has completed).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- 128-bit shared-memory access by 32-bit shared address --------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t a)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{ asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

}} // namespace ugemm::ptx
