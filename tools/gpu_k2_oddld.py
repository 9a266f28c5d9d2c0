#!/usr/bin/env python
"""Forced K2 on BASELINE config 3 with odd leading dimensions (K+5, K+3, N+7): the unguarded interior loop for 4-byte-aligned
operands (ANYLD instantiation) against the aligned-ld case.  JSON lines to gpurun_out/k2_oddld.jsonl."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugemm_b200 as u
u.sgemm_cuda_init()
M, N, K = 4095, 3001, 2047
with open(os.path.join(ROOT, "gpurun_out", "k2_oddld.jsonl"), "a") as log:
    for ta, tb in (("N", "T"), ("T", "N"), ("T", "T"), ("N", "N")):
        ar, ac = (M, K) if ta == "N" else (K, M)
        br, bc = (K, N) if tb == "N" else (N, K)
        for name, (pa, pb, pc) in (("odd", (5, 3, 7)), ("aligned", ((-ac) % 4, (-bc) % 4, (-N) % 4))):
            lda, ldb, ldc = ac + pa, bc + pb, N + pc
            dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
            dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
            avg, best = u.sgemm_cuda_time_dev("simt", 10, 2, "R", ta, tb, M, N, K, 1.5, dA, lda, dB, ldb, 0.5, dC, ldc)
            line = json.dumps({"shape": f"c3 {ta}{tb} {name} ld", "ms_avg": avg, "ms_min": best, "tflops": 2.0 * M * N * K / best / 1e9})
            print(line, flush=True); log.write(line + "\n")
            dA.free(); dB.free(); dC.free()
