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
        self.a = [ops.empty(p.mloc * p.kw) for _ in range(p.L)]
        self.b = [ops.empty(p.kw * p.nloc) for _ in range(p.L)]
        self.c = ops.empty(p.mloc * p.nloc)

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
