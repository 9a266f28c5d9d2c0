import sys, json
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {}
for (M, N, K) in ((4096, 32, 512), (4096, 8, 2048), (8192, 16, 4096), (65536, 8, 1024), (16384, 32, 2048), (32, 16384, 2048), (16, 65536, 512), (8, 200704, 1152), (2048, 32, 1024), (1024, 16, 8192), (100000, 12, 300), (30000, 20, 640)):
    dA, dB, dC = u.DeviceBuffer(M * K).fill_uniform(1), u.DeviceBuffer(K * N).fill_uniform(2), u.DeviceBuffer(M * N)
    row = {"mnk_log2": round(__import__("math").log2(M * N * K), 1)}
    for mode in ("3xtf32", "simt"):
        try:
            avg, best = u.sgemm_cuda_time_dev(mode, 10, 3, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
            row[mode] = round(avg * 1000, 1)
        except Exception as ex:
            row[mode] = None
            u.backend.lib().sgemm_cuda_clear_error()
    out[f"{M}x{N}x{K}"] = row
    dA.free(); dB.free(); dC.free()
print(json.dumps(out))
