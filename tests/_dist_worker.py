"""Worker for tests/test_dist_cpu.py: one rank of a gloo job running the sharded-GEMM schedule on CPU tensors.
The local GEMM is the ORACLE here (this file is test infrastructure); the product path is csrc/shard.cu)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402
from ugemm_b200 import backend as be  # noqa: E402
from ugemm_b200.dist import ShardedGemm, SlabPlan  # noqa: E402


class OracleOps:
    def empty(self, n):
        return torch.zeros(n, dtype=torch.float32)

    def fill_window(self, t, rows, cols, seed, offset, gld, lo, hi):
        be.fill_uniform_host_2d(rows, cols, seed, offset, gld, lo, hi, out=t.numpy())

    def gemm(self, M, N, K, A, lda, B, ldb, beta, Cm, ldc):
        out = Cm.numpy()
        O.oracle().oracle_sgemm_banded(1, b"R", b"N", b"N", M, N, K, 1.0, A.numpy(), lda, B.numpy(), ldb, beta, out, ldc)

    def sync(self):
        pass


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    M, N, K, L = (int(v) for v in sys.argv[1:5])
    out_dir = sys.argv[5]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = SlabPlan(world, rank, M, N, K, L=L)
    sg = ShardedGemm(plan, OracleOps(), dist)
    sg.generate_owned(seed_a=21, seed_b=22, lo=-0.5, hi=0.5)
    c = sg.run(distribute=True)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), c.numpy().reshape(plan.mloc, plan.nloc))
    # second run with resident panels must give the same block
    c2 = sg.run(distribute=False).numpy().copy()
    assert np.array_equal(c2.reshape(plan.mloc, plan.nloc), np.load(os.path.join(out_dir, f"c_{rank}.npy")))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
