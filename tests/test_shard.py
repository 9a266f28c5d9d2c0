"""Sharded SGEMM behind the C ABI (csrc/shard.cu, sgemm_cuda_shard_*): one process per GPU, NCCL broadcast / copy-engine pull.
CPU: the partition mirrors ugemm_b200/dist.py:SlabPlan, loud failure without init.  GPU: the driver at world = 1 in-process, and
world = 2 / 4 / 8 in as many processes as the box has GPUs (rendezvous through a file, no torch)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugemm_b200 as u  # noqa: E402
from ugemm_b200.dist import SlabPlan  # noqa: E402


def test_shard_plan_mirrors_slabplan():
    """sgemm_cuda_shard_plan / _owners (pure host arithmetic in C) against the Python plan the gloo tests exercise"""
    for world in (1, 2, 4, 8):
        for (M, N, K) in ((32768, 32768, 32768), (8192, 8192, 8192), (4096, 2048, 1024), (1024, 1024, 256), (64, 64, 32)):
            offs = {}
            for rank in range(world):
                p = SlabPlan(world, rank, M, N, K)
                q = u.Shard.plan(world, rank, M, N, K)
                assert (p.pr, p.pc, p.L, p.kw, p.mloc, p.nloc) == tuple(q[k] for k in ("pr", "pc", "L", "kw", "mloc", "nloc"))
                for t in range(p.L):
                    ao, bo, aoff, boff = u.Shard.owners(world, rank, M, N, K, t)
                    assert ao == p.a_owner(t) and bo == p.b_owner(t)
                    offs.setdefault(ao, set()).add(("a", rank // p.pc, t, aoff))
                    offs.setdefault(bo, set()).add(("b", rank % p.pc, t, boff))
            # inside one owner's allocation the slabs do not overlap and are dense
            p = SlabPlan(world, 0, M, N, K)
            for owner, items in offs.items():
                spans = sorted((off, off + (p.mloc * p.kw if kind == "a" else p.kw * p.nloc)) for kind, _, _, off in items)
                assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert u.last_error() is None


def test_shard_rejects_bad_shapes_and_missing_init():
    with pytest.raises(u.UgemmCudaError):
        u.Shard.plan(4, 0, 1001, 1000, 64)        # M does not divide the 2 x 2 grid
    u.backend.lib().sgemm_cuda_clear_error()
    import ctypes as C
    ms = C.c_float(0)
    assert u.backend.lib().sgemm_cuda_shard_run(1, 1, 0, C.byref(ms)) == 1      # not initialised: an error, not a crash
    assert "not initialised" in (u.last_error() or "")
    u.backend.lib().sgemm_cuda_clear_error()
    assert u.backend.lib().sgemm_cuda_shard_transport() == -1


def _spawn(world, M, N, K, transport, tmp_path):
    idfile = str(tmp_path / f"ncclid_{world}_{transport}")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_shard_worker.py"), str(r), str(world), idfile, str(M), str(N), str(K), str(transport)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    print("\n".join(o.strip()[-600:] for o in outs))
    assert all(pr.returncode == 0 for pr in procs), [pr.returncode for pr in procs]
    return outs


@pytest.mark.gpu
def test_shard_world1_in_process(tmp_path):
    _spawn(1, 768, 640, 2048, 0, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("transport", [0, 1])
def test_shard_multi_process(world, transport, tmp_path):
    if u.visible_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    outs = _spawn(world, 1024 * (2 if world > 1 else 1), 512 * (world // 2 if world > 2 else 1) * (2 if world > 2 else 1), 4096, transport, tmp_path)
    assert all("relerr" in o for o in outs)
