// shard.cu -- sharded SGEMM with ONE PROCESS PER GPU behind the C ABI (SURVEY section 8 e; BASELINE config 5).
//
// C = A * B (row-major NN, fp32) is cut into a pr x pc grid of C blocks; rank (i, j) owns block (i, j) and needs A row-panel i
// and B column-panel j.  K is not split across ranks: one exchange step, no reduction.  K is cut into L slabs; slab t of A
// panel i starts on ONE rank of grid row i (column t*pc/L), slab t of B panel j on ONE rank of grid column j (row t*pr/L) -- an
// owner-rooted placement, every rank holds 1/pc of its A panel and 1/pr of its B panel and no GPU's egress is the bottleneck.
// Two transports move the slabs over NVLink:
//   NCCL   ncclBroadcast from the owner inside the grid-row / grid-column communicator (ncclCommInitRank + ncclCommSplit), all
//          slabs issued up front on a communication stream; a few SMs are left to NCCL's CTAs while broadcasts are in flight
//          (a persistent K1 CTA fills an SM).
//   P2P    the owned slabs live in one allocation exported with cudaIpcGetMemHandle (handles all-gathered through NCCL);
//          receivers PULL every slab with a stream-ordered peer copy on the copy engines: no SM is taken from the GEMM and,
//          because a run never modifies owned slabs, a step needs no cross-process synchronisation at all.
// The local product of slab t (beta = 1 after the first) starts when ITS two transfers have landed, and the transfer of slab t
// of the next step only waits for the product that still reads the buffer: distribution overlaps the tensor work slab by slab.
//
// The reference owns one device and one queue (ocl.h:141-193); this file is the multi-process twin of sgemm_cuda_mgpu
// (backend.cu).  The host program brings RENDEZVOUS ONLY: rank 0 obtains a 128-byte id (sgemm_cuda_shard_unique_id) and hands it
// to every rank by whatever it has (torch.distributed in bench.py, MPI, a file).  NCCL is loaded with dlopen, so the library
// keeps no link-time dependency on it and single-GPU users never touch it.
// One sharded problem per process at a time, driven from one host thread at a time (the state below is a single object, like the
// backend's); every rank must make the same calls in the same order, as with any collective library.
#include "common.cuh"
#include "../../include/ugemm_cuda.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstdarg>
#include <cstring>
#include <ctime>
#include <vector>

namespace ugemm { void report_error(const char *msg); }

namespace {

void fail(const char *fmt, ...)
{
	char buf[400];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	ugemm::report_error(buf);
}
#define SH_CUDA(expr, what) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { fail("%s failed: %s", what, cudaGetErrorString(e__)); return 1; } } while (0)

// ---- NCCL through dlopen -------------------------------------------------------------------------------------------
struct Nccl {
	void *h = nullptr;
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&ncclCommInitRank) CommInitRank = nullptr;
	decltype(&ncclCommSplit) CommSplit = nullptr;
	decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclBroadcast) Broadcast = nullptr;
	decltype(&ncclAllGather) AllGather = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr;
	decltype(&ncclGroupStart) GroupStart = nullptr;
	decltype(&ncclGroupEnd) GroupEnd = nullptr;
	decltype(&ncclGetErrorString) GetErrorString = nullptr;
	bool load()
	{
		if (h) return true;
		// if the host program already loaded an NCCL (torch bundles one), the same soname resolves to it: one NCCL per process
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (h) break;
		}
		if (!h) { fail("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return false; }
#define SYM(n) n = reinterpret_cast<decltype(n)>(dlsym(h, "nccl" #n)); if (!n) { fail("libnccl lacks nccl" #n); h = nullptr; return false; }
		SYM(GetUniqueId) SYM(CommInitRank) SYM(CommSplit) SYM(CommDestroy) SYM(Broadcast) SYM(AllGather) SYM(AllReduce) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
		return true;
	}
} nccl;
#define SH_NCCL(expr, what) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) { fail("%s failed: %s", what, nccl.GetErrorString(r__)); return 1; } } while (0)

// ---- the partition (pure host arithmetic, mirrored by ugemm_b200/dist.py: SlabPlan and tested against it on CPU) ------
struct Plan { int world, rank, pr, pc, i, j, M, N, K, L, mloc, nloc, kw; };

void grid_shape(int world, int *pr, int *pc)
{
	if (world == 2) { *pr = 2; *pc = 1; return; }      // BASELINE config 5: 1x1, 2x1, 2x2, 2x4
	int r = 1;
	while ((r + 1) * (r + 1) <= world) r++;
	while (world % r) r--;
	*pr = r; *pc = world / r;
}
int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// the smallest multiple of lcm(pr, pc) whose slabs are <= 4096 wide: wider slabs mean fewer beta = 1 passes over C, narrower
// slabs less exposed first-slab transfer
int slab_count(int pr, int pc, int K)
{
	if (pr * pc == 1) return 1;
	int L = pr * pc / gcd_i(pr, pc);
	while (K / L > 4096 && K / (L * 2) >= 32 && K % (L * 2) == 0) L *= 2;
	return L;
}
int make_plan(int world, int rank, int M, int N, int K, Plan *p)
{
	if (world < 1 || rank < 0 || rank >= world || M < 1 || N < 1 || K < 1) { fail("shard plan: bad arguments (world %d rank %d M %d N %d K %d)", world, rank, M, N, K); return 1; }
	p->world = world; p->rank = rank; p->M = M; p->N = N; p->K = K;
	grid_shape(world, &p->pr, &p->pc);
	if (M % p->pr || N % p->pc) { fail("shard plan: M=%d / N=%d must divide the %d x %d grid", M, N, p->pr, p->pc); return 1; }
	p->i = rank / p->pc; p->j = rank % p->pc;
	p->L = slab_count(p->pr, p->pc, K);
	if (K % p->L || p->L % p->pr || p->L % p->pc) { fail("shard plan: K=%d must divide into L=%d slabs with L a multiple of pr and pc", K, p->L); return 1; }
	p->mloc = M / p->pr; p->nloc = N / p->pc; p->kw = K / p->L;
	return 0;
}
inline int a_owner(const Plan &p, int i, int t) { return i * p.pc + t * p.pc / p.L; }         // global rank holding slab t of A panel i
inline int b_owner(const Plan &p, int j, int t) { return (t * p.pr / p.L) * p.pc + j; }
// offset (floats) of an owned slab inside rank r's exported allocation: A and B slabs interleaved in slab order -- every rank
// can compute every other rank's layout, so only the 64-byte IPC handles travel
long long own_offset(const Plan &p, int r, bool is_b, int t, long long *total)
{
	const int ri = r / p.pc, rj = r % p.pc;
	const long long an = (long long)p.mloc * p.kw, bn = (long long)p.kw * p.nloc;
	long long off = 0, found = -1;
	for (int s = 0; s < p.L; s++) {
		if (a_owner(p, ri, s) == r) { if (!is_b && s == t) found = off; off += an; }
		if (b_owner(p, rj, s) == r) { if (is_b && s == t) found = off; off += bn; }
	}
	if (total) *total = off;
	return found;
}

// NCCL transport: SMs left to NCCL's CTAs while broadcasts may be in flight (a persistent K1 CTA fills an SM, and an NCCL kernel that
// finds no free SM waits for the whole product in front of it).  Every product but the last of a step runs on the reduced grid: the
// broadcasts of the NEXT step are issued right behind this step's products and need their SMs during this step's later slabs.
// [measured, 8 GPUs, NCCL_MAX_CTAS=4] reserving only for the first two products of a step -- with or without the host waiting for
// the step's last broadcast before it launches the rest -- serialises the next step's broadcasts behind full-grid products:
// 42 -> 82 ms per step (4 GPUs: 77 -> 108 ms).
constexpr int MAX_SLABS = 64, COMM_SMS = 8;
struct Shard {
	bool ready = false;
	Plan p{};
	int device = 0, sm_count = 148, transport = 0;       // transport actually in use: 0 NCCL broadcast, 1 copy-engine peer pull
	ncclComm_t comm = nullptr, row = nullptr, col = nullptr;
	cudaStream_t comp = nullptr, xfer = nullptr, up = nullptr, down = nullptr;
	float *own = nullptr, *recv = nullptr, *c = nullptr, *c_alt = nullptr, *scratch = nullptr;   // c_alt: second C block of the end-to-end path
	long long own_floats = 0;
	float *a[MAX_SLABS] = {nullptr}, *b[MAX_SLABS] = {nullptr};
	const float *a_src[MAX_SLABS] = {nullptr}, *b_src[MAX_SLABS] = {nullptr};   // P2P: where a non-owned slab is pulled from
	cudaEvent_t landed[MAX_SLABS] = {nullptr}, used[MAX_SLABS] = {nullptr}, uploaded[MAX_SLABS] = {nullptr};
	cudaEvent_t e0 = nullptr, e1 = nullptr, e_panel[8] = {nullptr}, c_down[2] = {nullptr, nullptr};
	std::vector<void *> mapped;
	float *h_own = nullptr, *h_c = nullptr;              // pinned host mirrors for the end-to-end path
} S;

int barrier_on(cudaStream_t st)
{
	if (S.p.world > 1) SH_NCCL(nccl.AllReduce(S.scratch, S.scratch, 1, ncclFloat, ncclSum, S.comm, st), "ncclAllReduce (barrier)");
	SH_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
	return 0;
}

// enqueue the distribution of slab t on S.xfer (it may start once `after`, if given, has completed on this device)
int distribute_slab(int t, cudaEvent_t after)
{
	const Plan &p = S.p;
	const size_t an = (size_t)p.mloc * p.kw, bn = (size_t)p.kw * p.nloc;
	if (after) SH_CUDA(cudaStreamWaitEvent(S.xfer, after, 0), "cudaStreamWaitEvent");
	if (S.transport == 1) {
		if (S.a_src[t]) SH_CUDA(cudaMemcpyAsync(S.a[t], S.a_src[t], an * 4, cudaMemcpyDefault, S.xfer), "peer pull (A slab)");
		if (S.b_src[t]) SH_CUDA(cudaMemcpyAsync(S.b[t], S.b_src[t], bn * 4, cudaMemcpyDefault, S.xfer), "peer pull (B slab)");
	} else {
		SH_NCCL(nccl.GroupStart(), "ncclGroupStart");
		if (p.pc > 1) SH_NCCL(nccl.Broadcast(S.a[t], S.a[t], an, ncclFloat, a_owner(p, p.i, t) % p.pc, S.row, S.xfer), "ncclBroadcast (A slab)");
		if (p.pr > 1) SH_NCCL(nccl.Broadcast(S.b[t], S.b[t], bn, ncclFloat, b_owner(p, p.j, t) / p.pc, S.col, S.xfer), "ncclBroadcast (B slab)");
		SH_NCCL(nccl.GroupEnd(), "ncclGroupEnd");
	}
	SH_CUDA(cudaEventRecord(S.landed[t], S.xfer), "cudaEventRecord");
	return 0;
}

void release_all()
{
	if (S.comp) cudaStreamSynchronize(S.comp);
	if (S.xfer) cudaStreamSynchronize(S.xfer);
	for (void *m : S.mapped) cudaIpcCloseMemHandle(m);
	S.mapped.clear();
	if (S.own) cudaFree(S.own);
	if (S.recv) cudaFree(S.recv);
	if (S.c) cudaFree(S.c);
	if (S.c_alt) cudaFree(S.c_alt);
	if (S.scratch) cudaFree(S.scratch);
	if (S.h_own) cudaFreeHost(S.h_own);
	if (S.h_c) cudaFreeHost(S.h_c);
	for (int t = 0; t < MAX_SLABS; t++) {
		if (S.landed[t]) cudaEventDestroy(S.landed[t]);
		if (S.used[t]) cudaEventDestroy(S.used[t]);
		if (S.uploaded[t]) cudaEventDestroy(S.uploaded[t]);
	}
	for (cudaEvent_t e : {S.e0, S.e1}) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : S.e_panel) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : S.c_down) if (e) cudaEventDestroy(e);
	for (cudaStream_t st : {S.comp, S.xfer, S.up, S.down}) if (st) cudaStreamDestroy(st);
	if (nccl.h) {
		if (S.row) nccl.CommDestroy(S.row);
		if (S.col) nccl.CommDestroy(S.col);
		if (S.comm) nccl.CommDestroy(S.comm);
	}
	cudaGetLastError();
	S = Shard();
}

} // namespace

extern "C" {

int sgemm_cuda_shard_plan(int world, int rank, int M, int N, int K, int *pr, int *pc, int *slabs, int *slab_width, int *block_rows, int *block_cols)
{
	Plan p;
	if (make_plan(world, rank, M, N, K, &p)) return 1;
	if (pr) *pr = p.pr;
	if (pc) *pc = p.pc;
	if (slabs) *slabs = p.L;
	if (slab_width) *slab_width = p.kw;
	if (block_rows) *block_rows = p.mloc;
	if (block_cols) *block_cols = p.nloc;
	return 0;
}

int sgemm_cuda_shard_owners(int world, int rank, int M, int N, int K, int t, int *a_owner_rank, int *b_owner_rank, long long *a_offset, long long *b_offset)
{
	Plan p;
	if (make_plan(world, rank, M, N, K, &p)) return 1;
	if (t < 0 || t >= p.L) { fail("shard owners: slab %d of %d", t, p.L); return 1; }
	const int ao = a_owner(p, p.i, t), bo = b_owner(p, p.j, t);
	if (a_owner_rank) *a_owner_rank = ao;
	if (b_owner_rank) *b_owner_rank = bo;
	if (a_offset) *a_offset = own_offset(p, ao, false, t, nullptr);
	if (b_offset) *b_offset = own_offset(p, bo, true, t, nullptr);
	return 0;
}

int sgemm_cuda_shard_unique_id(unsigned char *id128)
{
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	if (!id128) { fail("sgemm_cuda_shard_unique_id: NULL buffer"); return 1; }
	if (!nccl.load()) return 1;
	ncclUniqueId id;
	SH_NCCL(nccl.GetUniqueId(&id), "ncclGetUniqueId");
	memcpy(id128, &id, 128);
	return 0;
}

void sgemm_cuda_shard_finish(void) { release_all(); }

static int shard_init_body(int rank, int world, const unsigned char *id128, int M, int N, int K, int transport);

int sgemm_cuda_shard_init(int rank, int world, const unsigned char *id128, int M, int N, int K, int transport)
{
	if (S.ready) { fail("sgemm_cuda_shard_init: already initialised (call sgemm_cuda_shard_finish first)"); return 1; }
	const int rc = shard_init_body(rank, world, id128, M, N, K, transport);
	if (rc) release_all();          // nothing half-built stays behind: communicators, streams, events and allocations made so far are returned
	return rc;
}

static int shard_init_body(int rank, int world, const unsigned char *id128, int M, int N, int K, int transport)
{
	Plan p;
	if (make_plan(world, rank, M, N, K, &p)) return 1;
	if (p.L > MAX_SLABS) { fail("sgemm_cuda_shard_init: %d slabs (max %d)", p.L, MAX_SLABS); return 1; }
	int sm = 0, khz = 0;
	size_t hbm = 0;
	char name[64];
	if (ugemm_cuda_device_info(&sm, &khz, &hbm, name, (int)sizeof name)) return 1;     // initialises the backend lazily on the current device
	S.p = p;
	S.sm_count = sm;
	SH_CUDA(cudaGetDevice(&S.device), "cudaGetDevice");
	if (world > 1) {
		if (!id128) { fail("sgemm_cuda_shard_init: world > 1 needs the id of sgemm_cuda_shard_unique_id"); return 1; }
		if (!nccl.load()) return 1;
		ncclUniqueId id;
		memcpy(&id, id128, 128);
		SH_NCCL(nccl.CommInitRank(&S.comm, world, id, rank), "ncclCommInitRank");
		SH_NCCL(nccl.CommSplit(S.comm, p.i, p.j, &S.row, nullptr), "ncclCommSplit (grid row)");
		SH_NCCL(nccl.CommSplit(S.comm, p.pr + p.j, p.i, &S.col, nullptr), "ncclCommSplit (grid column)");
	}
	for (cudaStream_t *st : {&S.comp, &S.xfer, &S.up, &S.down}) SH_CUDA(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking), "cudaStreamCreate");
	for (int t = 0; t < p.L; t++) {
		SH_CUDA(cudaEventCreateWithFlags(&S.landed[t], cudaEventDisableTiming), "cudaEventCreate");
		SH_CUDA(cudaEventCreateWithFlags(&S.used[t], cudaEventDisableTiming), "cudaEventCreate");
		SH_CUDA(cudaEventCreateWithFlags(&S.uploaded[t], cudaEventDisableTiming), "cudaEventCreate");
	}
	SH_CUDA(cudaEventCreate(&S.e0), "cudaEventCreate");
	SH_CUDA(cudaEventCreate(&S.e1), "cudaEventCreate");
	for (cudaEvent_t &e : S.e_panel) SH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
	for (cudaEvent_t &e : S.c_down) SH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");

	// owned slabs in ONE allocation (exportable), received slabs in another, the C block, a scratch word for barriers
	const long long an = (long long)p.mloc * p.kw, bn = (long long)p.kw * p.nloc;
	long long own_total = 0, recv_total = 0;
	own_offset(p, rank, false, -1, &own_total);
	for (int t = 0; t < p.L; t++) {
		if (a_owner(p, p.i, t) != rank) recv_total += an;
		if (b_owner(p, p.j, t) != rank) recv_total += bn;
	}
	S.own_floats = own_total;
	SH_CUDA(cudaMalloc(&S.own, (size_t)(own_total > 0 ? own_total : 1) * 4), "cudaMalloc (owned slabs)");
	SH_CUDA(cudaMalloc(&S.recv, (size_t)(recv_total > 0 ? recv_total : 1) * 4), "cudaMalloc (received slabs)");
	SH_CUDA(cudaMalloc(&S.c, (size_t)p.mloc * p.nloc * 4), "cudaMalloc (C block)");
	SH_CUDA(cudaMalloc(&S.scratch, 256 + 64 * (size_t)world), "cudaMalloc (scratch)");
	SH_CUDA(cudaMemset(S.scratch, 0, 256 + 64 * (size_t)world), "cudaMemset");
	long long roff = 0;
	for (int t = 0; t < p.L; t++) {
		if (a_owner(p, p.i, t) == rank) S.a[t] = S.own + own_offset(p, rank, false, t, nullptr); else { S.a[t] = S.recv + roff; roff += an; }
		if (b_owner(p, p.j, t) == rank) S.b[t] = S.own + own_offset(p, rank, true, t, nullptr); else { S.b[t] = S.recv + roff; roff += bn; }
	}
	S.transport = 0;
	if (world > 1 && transport == 1) {
		// P2P: all-gather the 64-byte IPC handles through NCCL, map the allocations of the ranks of my grid row and column
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
		cudaIpcMemHandle_t mine;
		std::vector<cudaIpcMemHandle_t> all((size_t)world);
		unsigned char *stage = reinterpret_cast<unsigned char *>(S.scratch) + 256;
		int ok = cudaIpcGetMemHandle(&mine, S.own) == cudaSuccess ? 1 : 0;
		if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
		SH_CUDA(cudaMemcpyAsync(stage + 64 * (size_t)rank, &mine, 64, cudaMemcpyHostToDevice, S.comp), "cudaMemcpyAsync");
		SH_NCCL(nccl.AllGather(stage + 64 * (size_t)rank, stage, 64, ncclInt8, S.comm, S.comp), "ncclAllGather (IPC handles)");
		SH_CUDA(cudaMemcpyAsync(all.data(), stage, 64 * (size_t)world, cudaMemcpyDeviceToHost, S.comp), "cudaMemcpyAsync");
		SH_CUDA(cudaStreamSynchronize(S.comp), "cudaStreamSynchronize");
		std::vector<float *> base((size_t)world, nullptr);
		for (int r = 0; r < world && ok; r++) {
			if (r == rank || (r / p.pc != p.i && r % p.pc != p.j)) continue;
			void *m = nullptr;
			if (cudaIpcOpenMemHandle(&m, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
			S.mapped.push_back(m);
			base[(size_t)r] = static_cast<float *>(m);
		}
		// every rank must agree: one rank without a peer path sends everybody to the NCCL transport
		float flag = ok ? 0.f : 1.f;
		SH_CUDA(cudaMemcpyAsync(S.scratch, &flag, 4, cudaMemcpyHostToDevice, S.comp), "cudaMemcpyAsync");
		SH_NCCL(nccl.AllReduce(S.scratch, S.scratch, 1, ncclFloat, ncclSum, S.comm, S.comp), "ncclAllReduce");
		SH_CUDA(cudaMemcpyAsync(&flag, S.scratch, 4, cudaMemcpyDeviceToHost, S.comp), "cudaMemcpyAsync");
		SH_CUDA(cudaStreamSynchronize(S.comp), "cudaStreamSynchronize");
		if (flag == 0.f) {
			S.transport = 1;
			for (int t = 0; t < p.L; t++) {
				const int ao = a_owner(p, p.i, t), bo = b_owner(p, p.j, t);
				if (ao != rank) S.a_src[t] = base[(size_t)ao] + own_offset(p, ao, false, t, nullptr);
				if (bo != rank) S.b_src[t] = base[(size_t)bo] + own_offset(p, bo, true, t, nullptr);
			}
		} else {
			for (void *m : S.mapped) cudaIpcCloseMemHandle(m);
			S.mapped.clear();
		}
	}
	S.ready = true;
	return 0;
}

int sgemm_cuda_shard_transport(void) { return S.ready ? S.transport : -1; }

int sgemm_cuda_shard_block(float **d_c, int *rows, int *cols, int *row0, int *col0)
{
	if (!S.ready) { fail("sgemm_cuda_shard_block: not initialised"); return 1; }
	if (d_c) *d_c = S.c;
	if (rows) *rows = S.p.mloc;
	if (cols) *cols = S.p.nloc;
	if (row0) *row0 = S.p.i * S.p.mloc;
	if (col0) *col0 = S.p.j * S.p.nloc;
	return 0;
}

// every rank synthesises ONLY the slabs it owns, as windows of the global A (M x K) and B (K x N) streams (ugemm_fill_uniform_*)
int sgemm_cuda_shard_generate(unsigned long long seed_a, unsigned long long seed_b, float lo, float hi)
{
	if (!S.ready) { fail("sgemm_cuda_shard_generate: not initialised"); return 1; }
	SH_CUDA(cudaSetDevice(S.device), "cudaSetDevice");      // the calling thread may be a different one than at init
	const Plan &p = S.p;
	for (int t = 0; t < p.L; t++) {
		if (a_owner(p, p.i, t) == p.rank &&
		    ugemm_fill_uniform_dev_2d(S.a[t], (size_t)p.mloc, (size_t)p.kw, (size_t)p.kw, seed_a, (unsigned long long)p.i * p.mloc * p.K + (unsigned long long)t * p.kw, (unsigned long long)p.K, lo, hi, S.comp)) return 1;
		if (b_owner(p, p.j, t) == p.rank &&
		    ugemm_fill_uniform_dev_2d(S.b[t], (size_t)p.kw, (size_t)p.nloc, (size_t)p.nloc, seed_b, (unsigned long long)t * p.kw * p.N + (unsigned long long)p.j * p.nloc, (unsigned long long)p.N, lo, hi, S.comp)) return 1;
	}
	return barrier_on(S.comp);      // every owner's slabs are final before anyone pulls them
}

// `steps` steps (distribution of every slab, if asked, + the local products) after `warmup` untimed ones, bracketed by a barrier and
// a stream synchronisation on both sides; *ms_total = this rank's CUDA-event time of the timed steps (the caller takes the max)
int sgemm_cuda_shard_run(int distribute, int steps, int warmup, float *ms_total)
{
	if (!S.ready) { fail("sgemm_cuda_shard_run: not initialised"); return 1; }
	SH_CUDA(cudaSetDevice(S.device), "cudaSetDevice");      // the calling thread may be a different one than at init
	const Plan &p = S.p;
	const bool dist = distribute && p.world > 1;
	int rc = 0;
	for (int it = 0; it < warmup + steps && !rc; it++) {
		if (it == warmup) {
			if (barrier_on(S.xfer) || barrier_on(S.comp)) return 1;
			SH_CUDA(cudaEventRecord(S.e0, S.comp), "cudaEventRecord");
			SH_CUDA(cudaStreamWaitEvent(S.xfer, S.e0, 0), "cudaStreamWaitEvent");     // the distribution belongs to the timed region
		}
		if (dist) {
			// NCCL's CTAs need SMs and a persistent K1 CTA fills one: leave a few free while broadcasts are in flight
			if (S.transport == 0) sgemm_cuda_set_sm_limit(S.sm_count - COMM_SMS);
			for (int t = 0; t < p.L && !rc; t++) rc = distribute_slab(t, it > 0 ? S.used[t] : nullptr);
		}
		for (int t = 0; t < p.L && !rc; t++) {
			if (dist) SH_CUDA(cudaStreamWaitEvent(S.comp, S.landed[t], 0), "cudaStreamWaitEvent");
			if (dist && S.transport == 0 && t == p.L - 1) sgemm_cuda_set_sm_limit(0);     // nothing of this step left in flight behind the last slab
			rc = sgemm_cuda_dev(UGEMM_MODE_AUTO, S.comp, 'R', 'N', 'N', p.mloc, p.nloc, p.kw, 1.f, S.a[t], p.kw, S.b[t], p.nloc, t == 0 ? 0.f : 1.f, S.c, p.nloc);
			if (!rc) SH_CUDA(cudaEventRecord(S.used[t], S.comp), "cudaEventRecord");
		}
		sgemm_cuda_set_sm_limit(0);
	}
	if (rc) return 1;
	SH_CUDA(cudaEventRecord(S.e1, S.comp), "cudaEventRecord");
	if (barrier_on(S.comp) || barrier_on(S.xfer)) return 1;
	float ms = 0.f;
	SH_CUDA(cudaEventElapsedTime(&ms, S.e0, S.e1), "cudaEventElapsedTime");
	if (ms_total) *ms_total = ms;
	return 0;
}

// max (op = 0) or sum (op = 1) of one float over all ranks: lets a host program take "the max over ranks" without another library
int sgemm_cuda_shard_allreduce(float *value, int op)
{
	if (!S.ready || !value) { fail("sgemm_cuda_shard_allreduce: not initialised"); return 1; }
	SH_CUDA(cudaSetDevice(S.device), "cudaSetDevice");      // the calling thread may be a different one than at init
	if (S.p.world == 1) return 0;
	SH_CUDA(cudaMemcpyAsync(S.scratch + 1, value, 4, cudaMemcpyHostToDevice, S.comp), "cudaMemcpyAsync");
	SH_NCCL(nccl.AllReduce(S.scratch + 1, S.scratch + 1, 1, ncclFloat, op == 0 ? ncclMax : ncclSum, S.comm, S.comp), "ncclAllReduce");
	SH_CUDA(cudaMemcpyAsync(value, S.scratch + 1, 4, cudaMemcpyDeviceToHost, S.comp), "cudaMemcpyAsync");
	SH_CUDA(cudaStreamSynchronize(S.comp), "cudaStreamSynchronize");
	return 0;
}

// ---- end to end: owned slabs start in pinned HOST memory, the C block ends there -------------------------------------------
// Pinned host mirrors of the owned-slab allocation and of the C block (allocated on first use).
int sgemm_cuda_shard_host_buffers(float **h_own, long long *own_floats, float **h_c, long long *c_floats)
{
	if (!S.ready) { fail("sgemm_cuda_shard_host_buffers: not initialised"); return 1; }
	SH_CUDA(cudaSetDevice(S.device), "cudaSetDevice");      // the calling thread may be a different one than at init
	if (!S.h_own) SH_CUDA(cudaMallocHost(&S.h_own, (size_t)(S.own_floats > 0 ? S.own_floats : 1) * 4), "cudaMallocHost (owned slabs)");
	if (!S.h_c) SH_CUDA(cudaMallocHost(&S.h_c, (size_t)S.p.mloc * S.p.nloc * 4), "cudaMallocHost (C block)");
	if (!S.c_alt) SH_CUDA(cudaMalloc(&S.c_alt, (size_t)S.p.mloc * S.p.nloc * 4), "cudaMalloc (second C block)");
	if (h_own) *h_own = S.h_own;
	if (own_floats) *own_floats = S.own_floats;
	if (h_c) *h_c = S.h_c;
	if (c_floats) *c_floats = (long long)S.p.mloc * S.p.nloc;
	return 0;
}
// copy the device-resident owned slabs into the host mirror (benchmark set-up: the synthetic inputs are generated on the device)
int sgemm_cuda_shard_download_owned(void)
{
	if (sgemm_cuda_shard_host_buffers(nullptr, nullptr, nullptr, nullptr)) return 1;
	SH_CUDA(cudaMemcpyAsync(S.h_own, S.own, (size_t)S.own_floats * 4, cudaMemcpyDeviceToHost, S.comp), "cudaMemcpyAsync (D2H owned slabs)");
	return barrier_on(S.comp);
}
// One end-to-end step per iteration, PIPELINED: owned slab t goes up on the upload stream and is broadcast as soon as it has landed
// (NCCL: the broadcast is stream-ordered behind the owner's upload, so no cross-process event is needed), the product of slab t
// starts when its two broadcasts have landed, and the last slab's product runs in row panels whose finished C rows go down while
// the next panel is multiplied.  Host wall clock between two barriers; *ms_total = this rank's time for `steps` steps.
int sgemm_cuda_shard_run_host(int steps, int warmup, float *ms_total, long long *h2d_bytes_per_step, long long *d2h_bytes_per_step)
{
	if (sgemm_cuda_shard_host_buffers(nullptr, nullptr, nullptr, nullptr)) return 1;
	const Plan &p = S.p;
	const size_t an = (size_t)p.mloc * p.kw, bn = (size_t)p.kw * p.nloc;
	const int panels = p.mloc >= 4096 ? 4 : 1, prow = p.mloc / panels;
	timespec t0{}, t1{};
	int rc = 0;
	for (int it = 0; it < warmup + steps && !rc; it++) {
		if (it == warmup) {
			if (barrier_on(S.xfer) || barrier_on(S.comp)) return 1;
			clock_gettime(CLOCK_MONOTONIC, &t0);
		}
		// two C blocks take turns: the way down of a finished C (as long on the host link as half the uploads) overlaps the next
		// step's products instead of holding them up; a block is written again only when its previous content has gone down
		float *cbuf = (it & 1) ? S.c_alt : S.c;
		if (it >= 2) SH_CUDA(cudaStreamWaitEvent(S.comp, S.c_down[it & 1], 0), "cudaStreamWaitEvent");
		if (p.world > 1) sgemm_cuda_set_sm_limit(S.sm_count - COMM_SMS);
		for (int t = 0; t < p.L && !rc; t++) {
			// upload what this rank owns of slab t (the buffer may still be read by the previous step's product)
			const bool own_a = a_owner(p, p.i, t) == p.rank, own_b = b_owner(p, p.j, t) == p.rank;
			if ((own_a || own_b) && it > 0) SH_CUDA(cudaStreamWaitEvent(S.up, S.used[t], 0), "cudaStreamWaitEvent");
			if (own_a) SH_CUDA(cudaMemcpyAsync(S.a[t], S.h_own + (S.a[t] - S.own), an * 4, cudaMemcpyHostToDevice, S.up), "H2D (A slab)");
			if (own_b) SH_CUDA(cudaMemcpyAsync(S.b[t], S.h_own + (S.b[t] - S.own), bn * 4, cudaMemcpyHostToDevice, S.up), "H2D (B slab)");
			SH_CUDA(cudaEventRecord(S.uploaded[t], S.up), "cudaEventRecord");
			SH_CUDA(cudaStreamWaitEvent(S.xfer, S.uploaded[t], 0), "cudaStreamWaitEvent");
			if (p.world > 1) {
				const int keep = S.transport;
				S.transport = 0;                       // the end-to-end path always broadcasts (see above)
				rc = distribute_slab(t, it > 0 ? S.used[t] : nullptr);
				S.transport = keep;
			} else {
				SH_CUDA(cudaEventRecord(S.landed[t], S.xfer), "cudaEventRecord");
			}
		}
		for (int t = 0; t < p.L && !rc; t++) {
			SH_CUDA(cudaStreamWaitEvent(S.comp, S.landed[t], 0), "cudaStreamWaitEvent");
			if (t == p.L - 1) sgemm_cuda_set_sm_limit(0);
			if (t < p.L - 1 || panels == 1) {
				rc = sgemm_cuda_dev(UGEMM_MODE_AUTO, S.comp, 'R', 'N', 'N', p.mloc, p.nloc, p.kw, 1.f, S.a[t], p.kw, S.b[t], p.nloc, t == 0 ? 0.f : 1.f, cbuf, p.nloc);
				if (!rc && t == p.L - 1) {
					SH_CUDA(cudaEventRecord(S.e_panel[0], S.comp), "cudaEventRecord");
					SH_CUDA(cudaStreamWaitEvent(S.down, S.e_panel[0], 0), "cudaStreamWaitEvent");
					SH_CUDA(cudaMemcpyAsync(S.h_c, cbuf, (size_t)p.mloc * p.nloc * 4, cudaMemcpyDeviceToHost, S.down), "D2H (C block)");
				}
			} else {
				for (int q = 0; q < panels && !rc; q++) {
					const size_t r0 = (size_t)q * prow;
					rc = sgemm_cuda_dev(UGEMM_MODE_AUTO, S.comp, 'R', 'N', 'N', prow, p.nloc, p.kw, 1.f, S.a[t] + r0 * p.kw, p.kw, S.b[t], p.nloc, t == 0 ? 0.f : 1.f,
					                    cbuf + r0 * p.nloc, p.nloc);
					if (rc) break;
					SH_CUDA(cudaEventRecord(S.e_panel[q], S.comp), "cudaEventRecord");
					SH_CUDA(cudaStreamWaitEvent(S.down, S.e_panel[q], 0), "cudaStreamWaitEvent");
					SH_CUDA(cudaMemcpyAsync(S.h_c + r0 * p.nloc, cbuf + r0 * p.nloc, (size_t)prow * p.nloc * 4, cudaMemcpyDeviceToHost, S.down), "D2H (C panel)");
				}
			}
			if (!rc) SH_CUDA(cudaEventRecord(S.used[t], S.comp), "cudaEventRecord");
		}
		sgemm_cuda_set_sm_limit(0);
		if (!rc) SH_CUDA(cudaEventRecord(S.c_down[it & 1], S.down), "cudaEventRecord");
	}
	if (rc) return 1;
	SH_CUDA(cudaStreamSynchronize(S.down), "cudaStreamSynchronize");
	if (barrier_on(S.comp) || barrier_on(S.xfer)) return 1;
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (ms_total) *ms_total = (float)((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
	if (h2d_bytes_per_step) *h2d_bytes_per_step = S.own_floats * 4;
	if (d2h_bytes_per_step) *d2h_bytes_per_step = (long long)p.mloc * p.nloc * 4;
	return 0;
}

// The host-link floor of the end-to-end step on this box: the same bytes as sgemm_cuda_shard_run_host moves (owned slabs up, C block
// down, both directions at once on their own streams), no broadcast, no product.  Host wall clock between two barriers.
int sgemm_cuda_shard_copy_floor(int steps, float *ms_total)
{
	if (sgemm_cuda_shard_host_buffers(nullptr, nullptr, nullptr, nullptr)) return 1;
	const Plan &p = S.p;
	timespec t0{}, t1{};
	if (barrier_on(S.xfer) || barrier_on(S.comp)) return 1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int it = 0; it < steps; it++) {
		SH_CUDA(cudaMemcpyAsync(S.own, S.h_own, (size_t)S.own_floats * 4, cudaMemcpyHostToDevice, S.up), "H2D (owned slabs)");
		SH_CUDA(cudaMemcpyAsync(S.h_c, S.c, (size_t)p.mloc * p.nloc * 4, cudaMemcpyDeviceToHost, S.down), "D2H (C block)");
	}
	SH_CUDA(cudaStreamSynchronize(S.up), "cudaStreamSynchronize");
	SH_CUDA(cudaStreamSynchronize(S.down), "cudaStreamSynchronize");
	if (barrier_on(S.comp)) return 1;
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (ms_total) *ms_total = (float)((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
	return 0;
}

} // extern "C"
