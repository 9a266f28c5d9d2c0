"""One rank of the multi-process sharded SGEMM test (tests/test_shard.py spawns `world` of these; rendezvous through a file).
usage: python tests/_shard_worker.py rank world idfile M N K transport"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    M, N, K, transport = (int(x) for x in sys.argv[4:8])
    import ugemm_b200 as u
    import _oracle as O
    u.sgemm_cuda_init(rank)
    if rank == 0:
        uid = u.Shard.unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 60:
                raise SystemExit("rendezvous file never appeared")
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    sh = u.Shard(rank, world, uid if world > 1 else None, M, N, K, transport=transport)
    sh.generate(seed_a=11, seed_b=12, lo=-0.5, hi=0.5)
    ms = sh.run(True, 2, 1)
    c_ptr, rows, cols, r0, c0 = sh.block()
    got = np.empty(rows * cols, np.float32)
    u.backend.lib().ugemm_cuda_memcpy_d2h(got.ctypes.data, c_ptr, 4 * rows * cols)
    # the oracle on this rank's block: regenerated windows of the global streams
    A = u.fill_uniform_host_2d(rows, K, 11, r0 * K, K, lo=-0.5, hi=0.5)
    B = u.fill_uniform_host_2d(K, cols, 12, c0, N, lo=-0.5, hi=0.5)
    want = O.run14(O.oracle().oracle_sgemm_banded, "R", "N", "N", rows, cols, K, 1.0, A.ravel(), K, B.ravel(), cols, 0.0, np.zeros(rows * cols, np.float32), cols, threads=4)
    e = O.relerr("R", rows, cols, want, got, cols)
    # end to end from pinned host memory: same result within round-off
    sh.download_owned()
    sh.run_host(1, 0)
    e2e = sh.host_c().copy()
    # (not bit for bit: while NCCL broadcasts are in flight K1 runs on fewer SMs, which moves the stream-K cut points)
    e_host = O.relerr("R", rows, cols, want, e2e, cols)
    same = e_host <= 1e-5 and O.relerr("R", rows, cols, got, e2e, cols) <= 2e-6
    worst = sh.allreduce(e, "max")
    print(f"rank {rank}/{world} grid {sh.p['pr']}x{sh.p['pc']} L={sh.p['L']} transport={sh.transport} relerr={e:.3e} max={worst:.3e} e2e_ok={same} ms={ms:.3f}", flush=True)
    sh.finish()
    if not (e <= 1e-5 and same):
        raise SystemExit(f"rank {rank}: relerr {e:.3e}, e2e_ok={same}")


if __name__ == "__main__":
    main()
