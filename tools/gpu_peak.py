"""Measured dense-TF32 ceiling on this box by the same method MEASURED_PEAKS.json uses for bf16 (torch.matmul,
cuBLAS): fp32 inputs with allow_tf32, 8192^3, best of 10 (burst) and back to back for 3 s (sustained), plus bf16 and
plain fp32 (cuBLAS SGEMM) for context.  Library calls are used here ONLY to measure the roofline denominator."""
import json, sys, time
import torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
out = {}
for name, dt, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, True), ("fp32", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dt); b = torch.randn(n, n, device="cuda", dtype=dt)
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t0 = time.time(); iters = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20): torch.matmul(a, b)
        iters += 20; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[name] = {"burst_tflops": 2 * n**3 / best / 1e9, "sustained_tflops": 2 * n**3 * iters / e0.elapsed_time(e1) / 1e9}
print(json.dumps({"exp": "cublas_peaks", **out}))
