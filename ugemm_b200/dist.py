"""Sharded SGEMM: one large C = A*B partitioned across GPUs as a 2-D grid of C tiles.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the box, gloo on CPU for the host-logic
tests).  GPU (i, j) of a pr x pc grid owns  C[i*M/pr:(i+1)*M/pr, j*N/pc:(j+1)*N/pc]  and needs the A row-panel
i (M/pr x K) and the B column-panel j (K x N/pc).  K is NOT split across GPUs, so there is exactly one exchange
step and no reduction:

  * the K extent is cut into L slabs (L a multiple of lcm(pr, pc));
  * A row-panel i, slab t  (M/pr x K/L, tight)  initially lives on ONE rank of grid row i   (owner column t*pc//L),
    B column-panel j, slab t (K/L x N/pc, tight) initially lives on ONE rank of grid column j (owner row t*pr//L),
    i.e. every rank starts with 1/pc of its A panel and 1/pr of its B panel -- an owner-rooted placement, so no
    single GPU's NVLink egress is the bottleneck;
  * slab t is distributed with `broadcast` from its owner inside the grid-row (A) / grid-column (B) communicator,
    issued asynchronously for all slabs up front; the local GEMM of slab t (beta = 1 after the first slab) starts
    as soon as ITS two broadcasts have landed, so panel distribution overlaps the tensor-core work slab by slab.

Two transports move the slabs (both over NVLink, selected by the ops object):
  * "nccl": `torch.distributed.broadcast` from the owner inside the grid-row / grid-column communicator.  NCCL's
    CTAs need SMs, and a persistent K1 CTA fills an SM, so a few SMs are left free while broadcasts are in flight
    (sgemm_cuda_set_sm_limit) and NCCL is asked for few CTAs (NCCL_MAX_CTAS).
  * "p2p": every rank exports its owned-slab buffer as a CUDA IPC handle once; receivers PULL each slab with a
    stream-ordered peer copy, which runs on the copy engines and takes no SM from the GEMM.  NCCL is then only used
    for the handle exchange and for two one-element all-reduces per step that order the pulls across processes.

The local GEMM is the C-ABI device entry point (sgemm_cuda_dev); this module contains no arithmetic.  The
`ops` object isolates everything device-specific so the same schedule runs on CPU tensors under gloo in tests
(tests/test_dist_cpu.py injects its own CPU-checker ops object there; the product default is CudaOps).
"""
import math


def grid_shape(world):
    """pr x pc for 1/2/4/8 GPUs as in BASELINE config 5: 1x1, 2x1, 2x2, 2x4 (general: most square, pr <= pc)."""
    if world == 2:
        return 2, 1
    pr = int(math.isqrt(world))
    while world % pr:
        pr -= 1
    return pr, world // pr


def slab_count(pr, pc, K, target_kw=4096, min_kw=32):
    """Number of K slabs: the smallest multiple of lcm(pr, pc) whose slabs are <= target_kw wide (>= min_kw).
    Wider slabs mean fewer beta=1 passes over C; narrower slabs mean less exposed first-slab transfer."""
    if pr * pc == 1:
        return 1
    base = pr * pc // math.gcd(pr, pc)
    L = base
    while K // L > target_kw and K // (L * 2) >= min_kw and K % (L * 2) == 0:
        L *= 2
    return L


class SlabPlan:
    """Static description of who owns and who needs what.  Pure host logic (unit-tested on CPU)."""

    def __init__(self, world, rank, M, N, K, L=None):
        self.world, self.rank = world, rank
        self.pr, self.pc = grid_shape(world)
        if M % self.pr or N % self.pc:
            raise ValueError(f"M={M} / N={N} must divide the {self.pr}x{self.pc} grid")
        self.i, self.j = divmod(rank, self.pc)
        self.M, self.N, self.K = M, N, K
        self.L = L or slab_count(self.pr, self.pc, K)
        if K % self.L or self.L % self.pr or self.L % self.pc:
            raise ValueError(f"K={K} must divide into L={self.L} slabs with L a multiple of pr and pc")
        self.mloc, self.nloc, self.kw = M // self.pr, N // self.pc, K // self.L
        self.row_ranks = [self.i * self.pc + jj for jj in range(self.pc)]
        self.col_ranks = [ii * self.pc + self.j for ii in range(self.pr)]

    def a_owner(self, t):
        """global rank that initially holds slab t of this rank's A row-panel"""
        return self.i * self.pc + t * self.pc // self.L

    def b_owner(self, t):
        return (t * self.pr // self.L) * self.pc + self.j

    def a_window(self, t):
        """(row0, col0, rows, cols) of A row-panel i slab t inside the global M x K matrix"""
        return self.i * self.mloc, t * self.kw, self.mloc, self.kw

    def b_window(self, t):
        return t * self.kw, self.j * self.nloc, self.kw, self.nloc

    def c_window(self):
        return self.i * self.mloc, self.j * self.nloc, self.mloc, self.nloc

    def recv_bytes(self):
        """bytes this rank receives over the fabric in the distribution step"""
        a = sum(self.mloc * self.kw * 4 for t in range(self.L) if self.a_owner(t) != self.rank)
        b = sum(self.kw * self.nloc * 4 for t in range(self.L) if self.b_owner(t) != self.rank)
        return a + b


class CudaOps:
    """Device plumbing for the real thing: torch CUDA tensors for memory/streams/NCCL, C ABI for compute."""

    def __init__(self, mode="auto", comm_sms=8):
        import torch

        from . import backend
        self.torch, self.be, self.mode = torch, backend, mode
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.sm_count = backend.device_info()["sm_count"]
        self.comm_sms = comm_sms

    def reserve_for_comm(self, on):
        """While panel broadcasts are in flight leave `comm_sms` SMs to NCCL (a persistent K1 CTA fills an SM)."""
        self.be.set_sm_limit(self.sm_count - self.comm_sms if on else 0)

    def empty(self, n):
        return self.torch.empty(n, dtype=self.torch.float32, device=self.device)

    def _stream(self):
        # the C ABI reads a NULL stream as "the backend's own stream"; torch's default stream has handle 0, which
        # must be passed as cudaStreamLegacy (0x1) so the launch is ordered with torch's work and NCCL waits
        return self.torch.cuda.current_stream().cuda_stream or 1

    def fill_window(self, t, rows, cols, seed, offset, gld, lo, hi):
        self.be.fill_uniform_dev_2d(t.data_ptr(), rows, cols, cols, seed, offset, gld, lo, hi, stream=self._stream())

    def gemm(self, M, N, K, A, lda, B, ldb, beta, Cm, ldc):
        self.be.sgemm_cuda_dev(self.mode, self._stream(), "R", "N", "N", M, N, K, 1.0,
                               A.data_ptr(), lda, B.data_ptr(), ldb, beta, Cm.data_ptr(), ldc)

    def sync(self):
        self.torch.cuda.synchronize()


class RawBuf:
    """A window of a ugemm_cuda_malloc'ed allocation (what the p2p transport shares between processes)."""

    def __init__(self, ptr, n):
        self.ptr, self.n = ptr, n

    def data_ptr(self):
        return self.ptr


class CudaP2POps(CudaOps):
    """Copy-engine pull transport: buffers come from the C ABI (cudaMalloc) so they can be IPC-exported."""
    transport = "p2p"

    def __init__(self, mode="auto"):
        super().__init__(mode, comm_sms=0)
        self.copy_stream = self.torch.cuda.Stream()
        self._keep = []

    def alloc(self, n):
        buf = self.be.DeviceBuffer(n)
        self._keep.append(buf)
        return buf

    def empty(self, n):
        b = self.alloc(n)
        return RawBuf(b.ptr, n)


class ShardedGemm:
    """C_ij = A_i * B_j on this rank, with slab-wise owner-rooted panel broadcast overlapped with compute."""

    def __init__(self, plan, ops, dist=None):
        self.plan, self.ops = plan, ops
        if dist is None and plan.world > 1:
            import torch.distributed as dist
        self.dist = dist
        p = plan
        self.row_group = self.col_group = None
        if p.world > 1:
            # every rank must create every group, in the same order
            for ii in range(p.pr):
                g = dist.new_group([ii * p.pc + jj for jj in range(p.pc)])
                if ii == p.i:
                    self.row_group = g
            for jj in range(p.pc):
                g = dist.new_group([ii * p.pc + jj for ii in range(p.pr)])
                if jj == p.j:
                    self.col_group = g
        self.transport = getattr(ops, "transport", "nccl") if p.world > 1 else "local"
        if self.transport == "p2p":
            self._init_p2p()
        else:
            self.a = [ops.empty(p.mloc * p.kw) for _ in range(p.L)]
            self.b = [ops.empty(p.kw * p.nloc) for _ in range(p.L)]
        self.c = ops.empty(p.mloc * p.nloc)

    def _init_p2p(self):
        """Owned slabs live in ONE exported allocation; peers' allocations are mapped once."""
        p, ops, be = self.plan, self.ops, self.ops.be
        an, bn = p.mloc * p.kw, p.kw * p.nloc
        table, off = {}, 0
        for t in range(p.L):
            if p.a_owner(t) == p.rank:
                table[("a", t)] = off
                off += an
            if p.b_owner(t) == p.rank:
                table[("b", t)] = off
                off += bn
        own = ops.alloc(max(off, 1))
        gathered = [None] * p.world
        self.dist.all_gather_object(gathered, (be.ipc_export(own.ptr), table))
        peers = (set(p.row_ranks) | set(p.col_ranks)) - {p.rank}
        self._peer_base = {r: be.ipc_import(gathered[r][0]) for r in sorted(peers)}
        self.a, self.b, self._pull = [], [], []
        for t in range(p.L):
            for kind, owner, n, lst in (("a", p.a_owner(t), an, self.a), ("b", p.b_owner(t), bn, self.b)):
                if owner == p.rank:
                    lst.append(RawBuf(own.ptr + 4 * table[(kind, t)], n))
                else:
                    dst = ops.empty(n)
                    lst.append(dst)
                    self._pull.append((t, dst.ptr, self._peer_base[owner] + 4 * gathered[owner][1][(kind, t)], 4 * n))
        self._flag = ops.torch.zeros(1, device=ops.device)
        self._events = [ops.torch.cuda.Event() for _ in range(p.L)]

    def _run_p2p(self):
        p, ops, torch = self.plan, self.ops, self.ops.torch
        cur = torch.cuda.current_stream()
        self.dist.all_reduce(self._flag)            # stream-ordered: every owner's slabs are final before anyone pulls
        ops.copy_stream.wait_stream(cur)
        cs = ops.copy_stream.cuda_stream
        pulls = {}
        for t, dst, src, nbytes in self._pull:
            pulls.setdefault(t, []).append((dst, src, nbytes))
        for t in range(p.L):
            for dst, src, nbytes in pulls.get(t, ()):
                ops.be.memcpy_async(dst, src, nbytes, cs)
            self._events[t].record(ops.copy_stream)
        for t in range(p.L):
            cur.wait_event(self._events[t])
            ops.gemm(p.mloc, p.nloc, p.kw, self.a[t], p.kw, self.b[t], p.nloc, 0.0 if t == 0 else 1.0, self.c, p.nloc)
        self.dist.all_reduce(self._flag)            # nobody overwrites its owned slabs while a peer may still be pulling
        return self.c

    def generate_owned(self, seed_a, seed_b, lo=0.0, hi=1.0):
        """Each rank synthesises ONLY the slabs it owns, as windows of the global A (M x K) and B (K x N) streams."""
        p = self.plan
        for t in range(p.L):
            if p.a_owner(t) == p.rank:
                r0, c0, rows, cols = p.a_window(t)
                self.ops.fill_window(self.a[t], rows, cols, seed_a, r0 * p.K + c0, p.K, lo, hi)
            if p.b_owner(t) == p.rank:
                r0, c0, rows, cols = p.b_window(t)
                self.ops.fill_window(self.b[t], rows, cols, seed_b, r0 * p.N + c0, p.N, lo, hi)
        self.ops.sync()

    def run(self, distribute=True):
        """Distribution (optional: panels may already be resident from a previous run) + local GEMMs.
        Asynchronous w.r.t. the host on the GPU path; callers bracket it with events / synchronize."""
        p = self.plan
        if distribute and self.transport == "p2p":
            return self._run_p2p()
        works = [None] * p.L
        if distribute and p.world > 1:
            for t in range(p.L):
                wa = wb = None
                if p.pc > 1:
                    wa = self.dist.broadcast(self.a[t], src=p.a_owner(t), group=self.row_group, async_op=True)
                if p.pr > 1:
                    wb = self.dist.broadcast(self.b[t], src=p.b_owner(t), group=self.col_group, async_op=True)
                works[t] = (wa, wb)
        overlapped = distribute and p.world > 1 and hasattr(self.ops, "reserve_for_comm")
        if overlapped:
            self.ops.reserve_for_comm(True)
        for t in range(p.L):
            if works[t] is not None:
                for w in works[t]:
                    if w is not None:
                        w.wait()   # NCCL: makes the current stream wait for the collective; gloo: blocks
            if overlapped and t == p.L - 1:
                self.ops.reserve_for_comm(False)   # nothing left in flight behind the last slab
            self.ops.gemm(p.mloc, p.nloc, p.kw, self.a[t], p.kw, self.b[t], p.nloc, 0.0 if t == 0 else 1.0, self.c, p.nloc)
        if overlapped:
            self.ops.reserve_for_comm(False)
        return self.c
