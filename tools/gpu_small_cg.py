"""Small and mid-size squares on single CTAs (128 x 128 tiles) vs CTA pairs with the stream-K tail: where is the cross-over?"""
import sys, json, time
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
out = {}
for n in (512, 768, 1024, 1280, 1408, 1536):
    dA, dB, dC = u.DeviceBuffer(n * n).fill_uniform(1), u.DeviceBuffer(n * n).fill_uniform(2), u.DeviceBuffer(n * n)
    row = {}
    for cg in (0, 1, 2):
        u.set_k1_tuning(cta_group=cg)
        u.sync(); time.sleep(0.2)
        avg, best = u.sgemm_cuda_time_dev("3xtf32", 50, 5, "R", "N", "N", n, n, n, 1.0, dA, n, dB, n, 0.0, dC, n)
        row["auto" if cg == 0 else f"cg{cg}"] = [round(avg * 1e3, 1), round(best * 1e3, 1)]
    out[str(n)] = row
    for d in (dA, dB, dC): d.free()
print(json.dumps(out))
