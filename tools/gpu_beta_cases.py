import os, sys, json, time
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {"flags": int(os.environ.get("UGEMM_K1_FLAGS", "1"))}
cases = [("c3 NT b.5", 4095, 3001, 2047, "N", "T", 1.5, 0.5), ("c3 NT b0", 4095, 3001, 2047, "N", "T", 1.5, 0.0), ("c3 TN b.5", 4095, 3001, 2047, "T", "N", 1.5, 0.5),
         ("c4 NN b1", 200704, 256, 1152, "N", "N", 1.0, 1.0), ("c4 NN b0", 200704, 256, 1152, "N", "N", 1.0, 0.0),
         ("4096^3 b1", 4096, 4096, 4096, "N", "N", 1.0, 1.0), ("4096^3 b0", 4096, 4096, 4096, "N", "N", 1.0, 0.0)]
for (name, M, N, K, ta, tb, alpha, beta) in cases:
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    lda, ldb, ldc = (ac + 3) // 4 * 4, (bc + 3) // 4 * 4, (N + 3) // 4 * 4
    dA, dB, dC = u.DeviceBuffer(ar * lda), u.DeviceBuffer(br * ldb), u.DeviceBuffer(M * ldc)
    dA.fill_uniform(1); dB.fill_uniform(2); dC.fill_uniform(3)
    u.sync(); time.sleep(0.5)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 20, 3, "R", ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc)
    out[name] = [round(avg, 4), round(best, 4), round(2.0 * M * N * K / avg / 1e9, 1)]
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
