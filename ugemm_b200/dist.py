"""Host-side mirror of the sharded SGEMM's plan and schedule.

The product path is the C ABI: `csrc/shard.cu` (`sgemm_cuda_shard_*`, Python mirror `ugemm_b200.backend.Shard`) holds the
communicators (NCCL through dlopen), the buffers, streams and events, both transports (NCCL broadcast, copy-engine peer pull) and
the timing; `bench.py --gpus N` drives it and uses torch.distributed for the rendezvous of the 128-byte NCCL id only.

This module keeps what can be checked WITHOUT a GPU:
  * `SlabPlan` -- who owns and who needs what: one large C = A*B on a pr x pc grid of C tiles; rank (i, j) owns C block (i, j)
    and needs A row-panel i and B column-panel j; K is cut into L slabs (L a multiple of lcm(pr, pc)); slab t of A panel i starts
    on ONE rank of grid row i (column t*pc//L), slab t of B panel j on ONE rank of grid column j (row t*pr//L) -- an owner-rooted
    placement.  `sgemm_cuda_shard_plan` / `_owners` are the same arithmetic in C (tests/test_shard.py compares them).
  * `ShardedGemm` -- the schedule (all slab broadcasts issued up front, the product of slab t waits for ITS two broadcasts,
    beta = 1 after the first slab) written against an injected `ops` object, so that tests/test_dist_cpu.py runs it under gloo
    with world sizes 2 and 4 on CPU tensors and a CPU checker as the local GEMM.  It contains no arithmetic and no CUDA.
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


class ShardedGemm:
    """C_ij = A_i * B_j on this rank, with slab-wise owner-rooted panel broadcast overlapped with compute (the schedule of
    csrc/shard.cu's NCCL transport, against an injected ops object: `empty`, `fill_window`, `gemm`, `sync`)."""

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
