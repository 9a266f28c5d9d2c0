"""Timing cases for build variants of the TS kernel (UGEMM_CUDA_LIB selects the library): dense shapes on CTA pairs and on single CTAs."""
import os, sys, json, time
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"lib": os.path.basename(os.environ.get("UGEMM_CUDA_LIB", "libugemm_cuda.so"))}
cases = [("c3 NT b0", 4095, 3001, 2047, "N", "T", 1.5, 0.0, 0, 20), ("c3 NT b.5", 4095, 3001, 2047, "N", "T", 1.5, 0.5, 0, 20), ("c4 NN b0", 200704, 256, 1152, "N", "N", 1.0, 0.0, 0, 20),
         ("4096^3", 4096, 4096, 4096, "N", "N", 1.0, 0.0, 0, 20), ("8192^3", 8192, 8192, 8192, "N", "N", 1.0, 0.0, 0, 8),
         ("1024^3 cg1", 1024, 1024, 1024, "N", "N", 1.0, 0.0, 1, 50), ("1536^3 cg1", 1536, 1536, 1536, "N", "N", 1.0, 0.0, 1, 30),
         ("2560^3 cg1", 2560, 2560, 2560, "N", "N", 1.0, 0.0, 1, 20), ("200704x128x1152", 200704, 128, 1152, "N", "N", 1.0, 0.0, 0, 20)]
if os.environ.get("CASESET") == "2":      # the other operand layouts
    cases = [("c3 TN b.5", 4095, 3001, 2047, "T", "N", 1.5, 0.5, 0, 20), ("c3 TT b.5", 4095, 3001, 2047, "T", "T", 1.5, 0.5, 0, 20), ("c3 NN b0", 4095, 3001, 2047, "N", "N", 1.5, 0.0, 0, 20),
             ("4096^3 TN", 4096, 4096, 4096, "T", "N", 1.0, 0.0, 0, 20), ("c4 NN b1", 200704, 256, 1152, "N", "N", 1.0, 1.0, 0, 20), ("2048^3", 2048, 2048, 2048, "N", "N", 1.0, 0.0, 0, 30),
             ("16384x8192x2048 b1", 16384, 8192, 2048, "N", "N", 1.0, 1.0, 0, 6)]
for (name, M, N, K, ta, tb, alpha, beta, cg, iters) in cases:
    u.set_k1_tuning(cta_group=cg)
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    lda, ldb, ldc = (ac + 3) // 4 * 4, (bc + 3) // 4 * 4, (N + 3) // 4 * 4
    dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    u.sync(); time.sleep(0.5)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", iters, 3, "R", ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc)
    out[name] = [round(avg, 4), round(best, 4), round(2.0 * M * N * K / avg / 1e9, 1)]
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
