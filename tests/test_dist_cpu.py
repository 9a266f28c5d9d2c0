"""CPU suite for the multi-GPU path: the slab/ownership plan (pure host logic) and the broadcast schedule run
for real with world_size 2 and 4 over gloo, with the oracle standing in for the local GEMM."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle as O
from ugemm_b200 import backend as be
from ugemm_b200.dist import SlabPlan, grid_shape, slab_count

HERE = os.path.dirname(os.path.abspath(__file__))


def test_grid_shapes_follow_baseline_config5():
    assert [grid_shape(w) for w in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (2, 4)]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_partitions_everything_exactly_once(world):
    M = N = K = 32768
    plans = [SlabPlan(world, r, M, N, K) for r in range(world)]
    pr, pc = grid_shape(world)
    L = plans[0].L
    assert L == slab_count(pr, pc, K) and K % L == 0 and L % pr == 0 and L % pc == 0
    cover = np.zeros((pr, pc), int)
    for p in plans:
        r0, c0, rows, cols = p.c_window()
        assert (rows, cols) == (M // pr, N // pc) and r0 == p.i * rows and c0 == p.j * cols
        cover[p.i, p.j] += 1
        for t in range(L):
            # the owner of a slab is a member of the communicator it is broadcast in
            assert p.a_owner(t) in p.row_ranks and p.b_owner(t) in p.col_ranks
            # all members of a grid row agree on the owner of an A slab (same for columns / B)
            for q in plans:
                if q.i == p.i:
                    assert q.a_owner(t) == p.a_owner(t)
                if q.j == p.j:
                    assert q.b_owner(t) == p.b_owner(t)
    assert (cover == 1).all()
    # owner-rooted placement: every rank owns L/pc of its row's A slabs and L/pr of its column's B slabs
    for p in plans:
        assert sum(p.a_owner(t) == p.rank for t in range(L)) == L // pc
        assert sum(p.b_owner(t) == p.rank for t in range(L)) == L // pr
        want = (p.mloc * K * 4) * (pc - 1) // pc + (K * p.nloc * 4) * (pr - 1) // pr
        assert p.recv_bytes() == want
    if world == 8:  # 2.15 GB A panel + 1.07 GB B panel per GPU
        assert plans[0].mloc * K * 4 == 2147483648 and K * plans[0].nloc * 4 == 1073741824


def test_plan_rejects_indivisible_shapes():
    with pytest.raises(ValueError):
        SlabPlan(4, 0, 1001, 1000, 1024)
    with pytest.raises(ValueError):
        SlabPlan(8, 0, 1024, 1024, 1000, L=16)


@pytest.mark.parametrize("world,M,N,K,L", [(2, 96, 64, 140, 2), (2, 64, 48, 96, 4), (4, 64, 64, 128, 4)])
def test_sharded_schedule_over_gloo(world, M, N, K, L, tmp_path):
    port = 29500 + (os.getpid() + world * 7 + L) % 2000
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
    procs = []
    for r in range(world):
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(M), str(N), str(K), str(L), str(tmp_path)],
                                      env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out.decode(errors="ignore"))
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    # single-process truth from the same global streams
    A = be.fill_uniform_host(M * K, 21, -0.5, 0.5)
    B = be.fill_uniform_host(K * N, 22, -0.5, 0.5)
    want = O.run14(O.oracle().oracle_sgemm_banded, "R", "N", "N", M, N, K, 1.0, A, K, B, N, 0.0, np.zeros(M * N, np.float32), N, threads=1)
    want = want.reshape(M, N)
    pr, pc = grid_shape(world)
    for r in range(world):
        i, j = divmod(r, pc)
        blk = np.load(os.path.join(tmp_path, f"c_{r}.npy"))
        ref = want[i * M // pr:(i + 1) * M // pr, j * N // pc:(j + 1) * N // pc]
        e = np.linalg.norm(blk.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64))
        assert e <= 1e-6, (r, e)


def test_single_process_plan_invariants():
    """The partition sgemm_cuda_mgpu uses (C ABI, sgemm_cuda_mgpu_plan; host arithmetic, no GPU): blocks of a pr x pc grid tile C
    exactly once, K slabs tile K exactly once, and the sizes keep every local leading dimension TMA-eligible."""
    import ugemm_b200 as u
    assert u.sgemm_cuda_mgpu_plan(32768, 32768, 32768, 2, 4, 1) == (16384, 8192, 4, 8192)      # BASELINE config 5 on 8 GPUs
    assert u.sgemm_cuda_mgpu_plan(32768, 32768, 32768, 1, 1, 1) == (32768, 32768, 1, 32768)    # one GPU: nothing to overlap
    assert u.sgemm_cuda_mgpu_plan(32768, 32768, 32768, 2, 4, 0) == (16384, 8192, 1, 32768)     # overlap off: one slab
    rng = np.random.default_rng(5)
    for _ in range(2000):
        M, N, K = (int(x) for x in rng.integers(0, 40000, size=3))
        pr, pc = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        ov = int(rng.integers(0, 3))
        mb, nb, L, kw = u.sgemm_cuda_mgpu_plan(M, N, K, pr, pc, ov)
        assert mb % 4 == 0 and nb % 4 == 0
        assert pr * mb >= M and pc * nb >= N                       # the grid covers C ...
        assert mb <= -(-M // pr) + 3 and nb <= -(-N // pc) + 3     # ... with blocks no larger than the even split rounded up to 4
        rows = [max(0, min(mb, M - i * mb)) for i in range(pr)]
        cols = [max(0, min(nb, N - j * nb)) for j in range(pc)]
        assert sum(rows) == M and sum(cols) == N                   # every row and column of C belongs to exactly one block
        assert 1 <= L <= 64
        if K > 0:
            widths = [min(kw, K - t * kw) for t in range(L)]
            assert all(w > 0 for w in widths) and sum(widths) == K   # every k belongs to exactly one slab
        if L > 1:
            assert kw % 32 == 0 and ((ov and pr * pc > 1) or ov > 1)
        if ov == 0 or (pr * pc == 1 and ov < 2):
            assert (L, kw) == (1, K)
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_mgpu_plan(8, 8, 8, 0, 1)
