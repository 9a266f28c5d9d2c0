#!/usr/bin/env python
"""bench.py -- the headline SGEMM benchmark of BASELINE.json, measured on B200 through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5]

N = 1  workload c2: row-major NN SGEMM 8192^3 fp32, alpha=1 beta=0 (BASELINE configs[1]), auto dispatch -> K1
       (3xTF32 tcgen05).  A "step" is one GEMM over device-resident synthetic operands (805 MB > 126 MB L2, so
       nothing survives in L2 between steps).  `value` = TFLOP/s over K back-to-back steps, CUDA events.
N > 1  workload c5: 32768^3 sharded as a 2-D grid of C tiles (ugemm_b200/dist.py), launched by torchrun, one rank per
       GPU over NCCL.  A step = owner-rooted panel broadcast + local GEMMs (distribution INCLUDED); strong scaling.
`e2e`  the same metric through the reference-facing host-pointer call (sgemm_cuda with pinned HOST buffers:
       H2D of A and B, kernel, D2H of C inside the timed region), wall clock around blocking calls.
`roofline`     dominant kernel vs the measured tensor peak: MEASURED_PEAKS.json bf16 dense / 2 (TF32) / 3 (3 MMAs).
`cpu_baseline` the reference's own CPU SGEMM (oracle/_ref: unmodified sgemm_avx on all host cores over disjoint
       row slabs) on the whole workload when that fits the time budget, else on a bounded row-slab sample (a small
       sample flatters the reference: its C slab then stays in cache).  --impl reference prints that arm alone.
The oracle/ directory is only ever executed here as that CPU baseline, never as the thing measured for `value`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IDLE_GAP_S = 2.0      # idle time in front of every side-table shape (see other_shapes)

WORKLOADS = {
    "c1": dict(M=1024, N=1024, K=1024, desc="SGEMM row-major NN 1024x1024x1024 fp32 alpha=1 beta=0 (BASELINE configs[0]: the check_sgemm CPU case)"),
    "c2": dict(M=8192, N=8192, K=8192, desc="SGEMM row-major NN 8192x8192x8192 fp32 alpha=1 beta=0 (BASELINE configs[1])"),
    "c4": dict(M=200704, N=256, K=1152, desc="im2col-shaped SGEMM NN 200704x256x1152 fp32 (BASELINE configs[3])"),
    "c5": dict(M=32768, N=32768, K=32768, desc="SGEMM NN 32768^3 fp32 sharded as a 2-D C-tile grid (BASELINE configs[4])"),
}


def config_for(wl_name, world=1):
    """The `config` object of the JSON line -- built by ONE function for both arms (--impl ours / reference) so that the two lines
    describe the same workload with the same keys; everything implementation-specific goes to `details`."""
    wl = WORKLOADS[wl_name]
    M, N, K = wl["M"], wl["N"], wl["K"]
    mb = (M * K + K * N + M * N) * 4 / 1e6
    if wl_name == "c1":
        l2 = "inputs (A+B+C = %.1f MB) fit the 126 MB L2: a buffer larger than L2 is rewritten between timed steps (flush)" % mb
    elif world > 1:
        l2 = "inputs larger than L2 (per-GPU panels of a %.0f MB problem); no flush needed" % mb
    else:
        l2 = "inputs larger than L2 (A+B+C = %.0f MB vs 126 MB L2); no flush needed" % mb
    return {"workload": wl["desc"], "l2_policy": l2}


def host_cores():
    """Host threads this process may use.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers, which made
    the round-1 reference arm run on one core at N >= 2; the checker libraries take the thread count as an explicit argument."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return {"bf16_burst": float(p["bf16_tflops"]), "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
        except Exception:
            pass
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


REF_BUILD_NOTE = "gcc -O3 -march=x86-64-v3 -funroll-loops -ffp-contract=fast (the reference Makefile asks clang -Ofast -march=native; clang is absent and the .so must run on another host)"


def cpu_reference_arm(wl, seconds_target, steps=1, warmup=0, cores=None):
    """Reference CPU SGEMM on `cores` host threads (default: all this process may use) over workload `wl` (whole, or a bounded
    row-slab sample of it).  cores == 1 is sgemm_avx exactly as shipped (the reference has no threading).
    Returns (tflops, cores, kind, sample_description, ms_per_step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np

    import _oracle as O
    N, K = wl["N"], wl["K"]
    cores = cores or host_cores()
    r = O.ref()
    if r is not None:
        kind = "reference"
        if cores == 1:
            fn = lambda M, A, B, Cm: r.ref_sgemm_avx(b"R", b"N", b"N", M, N, K, 1.0, A, K, B, N, 0.0, Cm, N)
            what = "unmodified sgemm_avx (sgemm_avx256.h:392) as shipped, one core; built with " + REF_BUILD_NOTE
        else:
            fn = lambda M, A, B, Cm: r.ref_sgemm_avx_mt(cores, b"R", b"N", b"N", M, N, K, 1.0, A, K, B, N, 0.0, Cm, N)
            what = "unmodified sgemm_avx (sgemm_avx256.h:392) on disjoint row slabs; built with " + REF_BUILD_NOTE
    else:
        o = O.oracle()
        kind = "port"
        fn = lambda M, A, B, Cm: o.oracle_sgemm_banded(cores, b"R", b"N", b"N", M, N, K, 1.0, A, K, B, N, 0.0, Cm, N)
        what = "oracle port of sgemm_avx's 35-band order"
    B = O.fill_uniform(K * N, 2)
    # Size the sample.  sgemm_avx re-reads and re-writes its C slab once per 35-wide K band (sgemm_avx256.h:316-390), so its rate
    # depends on the slab height per thread: 64 rows per thread stay in the core's L2 (1.0-1.1 TFLOP/s on 16 threads), the
    # 512 rows per thread of the real 8192-row workload do not (0.45 TFLOP/s).  A small sample therefore FLATTERS the reference;
    # the whole workload is run whenever it fits the time budget, and only otherwise a bounded slab.
    m0 = min(64 * cores, wl["M"])
    A = O.fill_uniform(m0 * K, 1)
    Cm = np.zeros(m0 * N, np.float32)
    fn(m0, A, B, Cm)                       # first call pays thread start-up and page faults
    t = time.perf_counter(); fn(m0, A, B, Cm); dt = max(time.perf_counter() - t, 1e-4)
    slow = 2.5                             # allowance for the cache effect above when extrapolating from the small slab
    if wl["M"] / m0 * dt * slow <= seconds_target:
        rows = wl["M"]
    else:
        rows = int(max(m0, seconds_target / (dt * slow) * m0))
        rows -= rows % (2 * cores)
        rows = min(max(rows, m0), wl["M"])
    A = O.fill_uniform(rows * K, 1)
    Cm = np.zeros(rows * N, np.float32)
    for _ in range(warmup):
        fn(rows, A, B, Cm)
    t = time.perf_counter()
    for _ in range(steps):
        fn(rows, A, B, Cm)
    dt = (time.perf_counter() - t) / steps
    tflops = 2.0 * rows * N * K / dt / 1e12
    part = "the whole workload" if rows == wl["M"] else f"rows 0..{rows - 1} of C"
    sample = f"{what}; {part} ({rows}x{N}x{K} of the {wl['M']}x{N}x{K} workload, {rows // cores} rows per thread), {cores} threads"
    return tflops, cores, kind, sample, dt * 1e3


def run_reference_impl(args, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[wl_name]
    tflops, cores, kind, sample, ms = cpu_reference_arm(wl, seconds_target=min(3.0, max(0.3, 150.0 / (max(args.steps, 1) + min(args.warmup, 2)))),
                                                       steps=max(args.steps, 1), warmup=min(args.warmup, 2), cores=1 if wl_name == "c1" else None)
    line = {"impl": "reference", "metric": "SGEMM TFLOP/s", "value": tflops, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(wl_name, max(args.gpus, 1)),
            "details": {"note": "reference CPU path on the GPU box's host cores; each step is the whole workload when that fits the time budget, else a bounded row-slab sample (see cpu_baseline.sample)",
                        "host_cores_used": cores, "OMP_NUM_THREADS_env": os.environ.get("OMP_NUM_THREADS")},
            "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def k1_traffic_bytes(wl_name, live=True):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on this workload.  Measured live when ncu
    is on the box: this script re-runs itself (`--traffic-child`: device-resident operands, warm-ups, a few launches) under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` after every timed region has finished, so no
    reported time is taken under the profiler.  Falls back to the committed capture (profiles/traffic.json).
    Returns (bytes or None, source)."""
    if live and wl_name in ("c1", "c2", "c4"):
        try:
            cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
                   "-k", "regex:k1ts_kernel|k1_3xtf32|k2_simt", "-s", "3", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__),
                   "--traffic-child", "--workload", wl_name]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            import csv
            import io
            tot, seen = 0.0, 0
            for row in csv.reader(io.StringIO(res.stdout)):
                if len(row) >= 3 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(row[-1].replace(",", ""))
                    seen += 1
            if seen == 2 and tot > 0:
                return int(tot), "live: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one launch, run by bench.py after the timed regions"
        except Exception:  # noqa: BLE001
            pass
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        j = json.load(open(path))
        return j.get(wl_name), "committed capture: " + j.get("source", path)
    except Exception:
        return None, "unavailable"


def run_traffic_child(wl_name):
    """Body of the ncu child process: nothing is timed or printed here."""
    import ugemm_b200 as u
    wl = WORKLOADS[wl_name]
    M, N, K = wl["M"], wl["N"], wl["K"]
    u.sgemm_cuda_init(0)
    dA, dB, dC = u.DeviceBuffer(M * K).fill_uniform(1), u.DeviceBuffer(K * N).fill_uniform(2), u.DeviceBuffer(M * N).fill_uniform(3)
    u.sgemm_cuda_time_dev("auto", 3, 3, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
    u.sync()


def other_shapes(u, peaks, info):
    """The other named BASELINE shapes at 1 GPU (north_star: every named shape in TFLOP/s and as a fraction of the relevant
    roofline).  Reported beside the headline, never part of `value`.  Each: 3 warm-ups, 10 launches, CUDA events, best and mean."""
    tensor_peak = peaks["bf16_burst"] / 6.0
    ffma_peak = info["sm_count"] * 128 * 2 * info["sm_clock_khz"] * 1e3 / 1e12      # SMs x 128 lanes x 2 flop x max clock
    out = []

    def run(name, mode, ta, tb, M, N, K, alpha, beta, lda, ldb, ldc, bound):
        ar, br = (M if ta == "N" else K), (K if tb == "N" else N)
        dA, dB, dC = u.DeviceBuffer(ar * lda).fill_uniform(1), u.DeviceBuffer(br * ldb).fill_uniform(2), u.DeviceBuffer(M * ldc).fill_uniform(3)
        # Each shape starts after an idle gap: the board leaves the headline run power-capped (clocks 10-15 % down for a while), and a
        # sub-millisecond shape measured straight after it reports the previous workload's throttle state, not its own
        # (4095x3001x2047: 0.29 ms right after twenty 8192^3 launches, 0.24 ms from idle, same binary, same box).
        u.sync()
        time.sleep(IDLE_GAP_S)
        avg, best = u.sgemm_cuda_time_dev(mode, 10, 3, "R", ta, tb, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc)
        kern = u.last_kernel() + (" (operands repacked for TMA)" if u.last_repacked() else "")
        flops = 2.0 * M * N * K
        nbytes = 4.0 * (M * K + K * N + M * N * (2 if beta != 0 else 1))
        peak = tensor_peak if u.last_kernel() == "3xtf32" else ffma_peak
        rec = {"shape": name, "mode": mode, "kernel": kern, "ms_avg": avg, "ms_min": best, "tflops": flops / avg / 1e9,
               "roofline_bound": "tensor (3xTF32)" if u.last_kernel() == "3xtf32" else "fp32 FFMA", "roofline_peak_tflops": peak,
               "roofline_frac": flops / avg / 1e9 / peak, "algorithmic_gbs": nbytes / avg / 1e6,
               "hbm_frac": nbytes / avg / 1e6 / peaks["hbm_gbs"], "idle_gap_s": IDLE_GAP_S}
        if bound:
            rec["note"] = bound
        out.append(rec)
        for b in (dA, dB, dC):
            b.free()

    run("c2 8192^3 NN, K2 plain-fp32 mode", "simt", "N", "N", 8192, 8192, 8192, 1.0, 0.0, 8192, 8192, 8192, None)
    for ta, tb in (("N", "T"), ("T", "N"), ("T", "T")):
        M, N, K = 4095, 3001, 2047
        lda = (K if ta == "N" else M) + 1   # 2048 / 4096: multiples of 4 -> TMA-eligible
        ldb = (N if tb == "N" else K) + (3 if tb == "N" else 1)
        run(f"c3 4095x3001x2047 {ta}{tb} alpha=1.5 beta=0.5, ld padded to a multiple of 4", "auto", ta, tb, M, N, K, 1.5, 0.5, lda, ldb, N + 3, None)
    run("c3 4095x3001x2047 NT alpha=1.5 beta=0.5, odd ld (K+5, K+3, N+7)", "auto", "N", "T", 4095, 3001, 2047, 1.5, 0.5, 2047 + 5, 2047 + 3, 3001 + 7,
        "TMA-ineligible layout: auto rule (2) repacks, forced simt runs K2 on it directly")
    run("c4 200704x256x1152 NN (im2col shape)", "auto", "N", "N", 200704, 256, 1152, 1.0, 0.0, 1152, 256, 256,
        "AI 105 flop/B: above the 3xTF32 ridge, HBM time is ~40% of the tensor time (SURVEY 8d)")
    run("c4 200704x256x1152 NN beta=1 (accumulate)", "auto", "N", "N", 200704, 256, 1152, 1.0, 1.0, 1152, 256, 256, None)
    out += widened_rows(u, peaks, info, tensor_peak)
    return out


def widened_rows(u, peaks, info, tensor_peak):
    """SURVEY section 8(f) rows at 1 GPU: the fused convolution that produces config 4's GEMM, SAXPY / SGEMV against the measured
    HBM bandwidth, DGEMM against the FP64 pipe.  Wall clock around 10 back-to-back launches bracketed by device syncs."""
    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        u.sync()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        u.sync()
        return (time.perf_counter() - t0) / iters * 1e3

    rows = []
    nimg, ich, h, w, k, pad, ch = 64, 128, 56, 56, 3, 1, 256
    dx, dw = u.DeviceBuffer(nimg * ich * h * w).fill_uniform(1, -0.5, 0.5), u.DeviceBuffer(ch * ich * k * k).fill_uniform(2, -0.5, 0.5)
    db, dout = u.DeviceBuffer(ch).fill_uniform(3, -0.5, 0.5), u.DeviceBuffer(nimg * ch * h * w)
    ms = timed(lambda: u.convolution_cuda_batched_dev("auto", None, dx, nimg, ich, w, h, dw, k, pad, 1, dout, ch, db, 0.1, None))
    flops = 2.0 * nimg * ch * h * w * ich * k * k
    rows.append({"shape": "convolution 64 x (128x56x56) -> 256 filters 3x3 pad 1 + bias + LeakyReLU (the layer behind c4), implicit GEMM",
                 "kernel": "3xtf32 CONV (4-D TMA gather)" if u.last_conv_fused() else "im2col + GEMM", "ms_avg": ms, "tflops": flops / ms / 1e9,
                 "roofline_bound": "tensor (3xTF32)", "roofline_peak_tflops": tensor_peak, "roofline_frac": flops / ms / 1e9 / tensor_peak,
                 "note": "time includes the channels-last staging pass and the weight repack; the 925 MB column matrix is never built"})
    for b in (dx, dw, db, dout):
        b.free()
    n = 1 << 28
    dx, dy = u.DeviceBuffer(n).fill_uniform(1), u.DeviceBuffer(n).fill_uniform(2)
    ms = timed(lambda: u.saxpy_cuda_dev(None, n, 0.5, dx, 1, dy, 1))
    rows.append({"shape": "saxpy n=2^28", "kernel": "saxpy_vec", "ms_avg": ms, "algorithmic_gbs": 12.0 * n / ms / 1e6, "roofline_bound": "hbm",
                 "roofline_peak_gbs": peaks["hbm_gbs"], "roofline_frac": 12.0 * n / ms / 1e6 / peaks["hbm_gbs"]})
    dx.free(); dy.free()
    M = N = 16384
    dA, dx, dy = u.DeviceBuffer(M * N).fill_uniform(3), u.DeviceBuffer(N).fill_uniform(4, -0.5, 0.5), u.DeviceBuffer(M)
    for trans in ("T", "N"):
        ms = timed(lambda: u.sgemv_cuda_dev(None, trans, M, N, 1.0, dA, M, dx, 1, 0.0, dy, 1))
        rows.append({"shape": f"sgemv '{trans}' 16384x16384", "kernel": "sgemv_rows" if trans == "T" else "sgemv_cols", "ms_avg": ms,
                     "algorithmic_gbs": 4.0 * M * N / ms / 1e6, "roofline_bound": "hbm", "roofline_peak_gbs": peaks["hbm_gbs"],
                     "roofline_frac": 4.0 * M * N / ms / 1e6 / peaks["hbm_gbs"]})
    for b in (dA, dx, dy):
        b.free()
    n = 4096
    dA, dB, dC = u.DeviceBuffer(2 * n * n).fill_uniform(5), u.DeviceBuffer(2 * n * n).fill_uniform(6), u.DeviceBuffer(2 * n * n)
    # the fp32 fill makes each double a pair of random fp32 words: finite, in [2^-127, 2) -- fine for a throughput run
    avg, best = u.dgemm_cuda_time_dev(10, 3, "R", "N", "N", n, n, n, 1.0, dA, n, dB, n, 0.0, dC, n)
    fp64_peak = info["sm_count"] * 64 * 2 * info["sm_clock_khz"] * 1e3 / 1e12
    rows.append({"shape": "dgemm 4096^3 NN", "kernel": "K4 DMMA (mma.sync.m8n8k4.f64)", "ms_avg": avg, "ms_min": best, "tflops": 2.0 * n ** 3 / avg / 1e9,
                 "roofline_bound": "fp64 pipe", "roofline_peak_tflops": fp64_peak, "roofline_frac": 2.0 * n ** 3 / avg / 1e9 / fp64_peak})
    for b in (dA, dB, dC):
        b.free()
    return rows


def run_single(args, wl_name):
    import numpy as np

    import ugemm_b200 as u
    wl = WORKLOADS[wl_name]
    M, N, K = wl["M"], wl["N"], wl["K"]
    u.sgemm_cuda_init(0)
    info = u.device_info()
    dA = u.DeviceBuffer(M * K).fill_uniform(1)
    dB = u.DeviceBuffer(K * N).fill_uniform(2)
    dC = u.DeviceBuffer(M * N).fill_uniform(3)
    u.sync()
    flops = 2.0 * M * N * K
    sampler = ClockSampler(0)
    sampler.start()
    l0 = u.launch_count()
    if (M * K + K * N + M * N) * 4 > 126e6:
        avg_ms, min_ms, total_ms = u.sgemm_cuda_time_dev("auto", args.steps, args.warmup, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N, total=True)
        launches = u.launch_count() - l0 - args.warmup
    else:
        # the operands fit the L2: rewrite a 256 MB buffer (larger than L2) before every timed launch, time each launch alone
        flush = u.DeviceBuffer(64 << 20)
        u.sgemm_cuda_time_dev("auto", 1, args.warmup, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
        l0 = u.launch_count()
        per = []
        for _ in range(args.steps):
            flush.fill_uniform(9)
            per.append(u.sgemm_cuda_time_dev("auto", 1, 0, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)[0])
        launches = u.launch_count() - l0 - args.steps          # the flush fills are not GEMM launches
        flush.free()
        avg_ms, min_ms, total_ms = sum(per) / len(per), min(per), sum(per)
    kernel = u.last_kernel()
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = flops / ms_per_step / 1e9

    # ---- e2e: the drop-in host-pointer call on pinned host buffers (H2D + kernel + D2H inside the timed region)
    import ctypes as C
    L = u.lib()
    hA, hB, hC = (L.ugemm_cuda_malloc_host(n * 4) for n in (M * K, K * N, M * N))
    if not (hA and hB and hC):
        u.check()
    L.ugemm_cuda_memcpy_d2h(hA, C.c_void_p(dA.ptr), M * K * 4)
    L.ugemm_cuda_memcpy_d2h(hB, C.c_void_p(dB.ptr), K * N * 4)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(min(args.warmup, 2)):
        u.sgemm_cuda("R", "N", "N", M, N, K, 1.0, hA, K, hB, N, 0.0, hC, N)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        u.sgemm_cuda("R", "N", "N", M, N, K, 1.0, hA, K, hB, N, 0.0, hC, N)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    # cheap integrity check of the e2e result against the device-resident run (same inputs, same kernel).  Not bit for bit: the host
    # path multiplies row panels, whose tiles fall into other waves of the persistent grid than the whole problem's, and the order in
    # which a tile walks K (serpentine) and the stream-K cut points depend on that -- two roundings of the same product.
    res_host = np.ctypeslib.as_array(C.cast(hC, C.POINTER(C.c_float)), shape=(M * N,))
    res_dev = dC.download(4096)
    diff = float(np.linalg.norm(res_host[:4096].astype(np.float64) - res_dev) / np.linalg.norm(res_dev.astype(np.float64)))
    assert diff <= 2e-6, f"e2e result differs from the device-resident result: relative difference {diff:.3e}"
    for h in (hA, hB, hC):
        L.ugemm_cuda_free_host(h)

    peaks = measured_peaks()
    side = wl_name == "c2" and not args.no_shapes
    sustained = c5_1gpu = config1 = None
    shapes = other_shapes(u, peaks, info) if side else None
    if side:
        # the same kernel back to back for >= 4 s: the power-capped rate, against the SUSTAINED measured peak
        iters = int(min(4096, max(args.steps, 4300.0 / avg_ms)))
        smp = ClockSampler(0)
        smp.start()
        s_avg, s_min, s_total = u.sgemm_cuda_time_dev("auto", iters, 1, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N, total=True)
        s_clk = smp.stop()
        s_tf = flops * iters / s_total / 1e9
        sustained = {"tflops": s_tf, "seconds": s_total / 1e3, "launches": iters, "ms_avg": s_avg, "peak": peaks["bf16_sustained"] / 6.0,
                     "frac": s_tf / (peaks["bf16_sustained"] / 6.0), "nominal_frac": s_tf / 375.0,
                     "peak_basis": f"{peaks['source']} bf16 dense sustained {peaks['bf16_sustained']:.1f} TFLOP/s / 6", "clocks": s_clk}
    for b in (dA, dB, dC):
        b.free()
    if side:
        # BASELINE config 5 on ONE GPU: the same-problem anchor of the 1/2/4/8-GPU scaling curve (operands 3 x 4.29 GB, device-generated)
        n5 = 32768
        d5 = [u.DeviceBuffer(n5 * n5).fill_uniform(s) for s in (1, 2)] + [u.DeviceBuffer(n5 * n5)]
        a5, m5 = u.sgemm_cuda_time_dev("auto", 2, 1, "R", "N", "N", n5, n5, n5, 1.0, d5[0], n5, d5[1], n5, 0.0, d5[2], n5)
        c5_1gpu = {"workload": WORKLOADS["c5"]["desc"], "ms_avg": a5, "ms_min": m5, "tflops": 2.0 * n5 ** 3 / a5 / 1e9, "launches": 2}
        for b in d5:
            b.free()
        # BASELINE config 1 on this box's host: sgemm_avx as shipped (ONE core, the reference has no threading) at 1024^3,
        # beside K1 on the same problem (L2 flushed between launches)
        c1_tf, _, c1_kind, c1_sample, c1_ms = cpu_reference_arm(WORKLOADS["c1"], seconds_target=2.0, steps=5, warmup=1, cores=1)
        config1 = {"workload": WORKLOADS["c1"]["desc"], "cpu_1core_gflops": c1_tf * 1e3, "cpu_ms": c1_ms, "kind": c1_kind, "sample": c1_sample}
        d1 = [u.DeviceBuffer(1 << 20).fill_uniform(s) for s in (1, 2, 3)]
        flush = u.DeviceBuffer(64 << 20)
        u.sgemm_cuda_time_dev("auto", 1, 3, "R", "N", "N", 1024, 1024, 1024, 1.0, d1[0], 1024, d1[1], 1024, 0.0, d1[2], 1024)
        per = []
        for _ in range(10):
            flush.fill_uniform(9)
            per.append(u.sgemm_cuda_time_dev("auto", 1, 0, "R", "N", "N", 1024, 1024, 1024, 1.0, d1[0], 1024, d1[1], 1024, 0.0, d1[2], 1024)[0])
        config1.update({"gpu_ms_avg_l2_cold": sum(per) / len(per), "gpu_tflops_l2_cold": 2.0 * 1024 ** 3 / (sum(per) / len(per)) / 1e9, "gpu_kernel": u.last_kernel()})
        for b in d1 + [flush]:
            b.free()
    # a kernel timed alone over a short burst of steps -> burst peak; TF32 dense = bf16 dense / 2; 3 MMAs per product
    peak = peaks["bf16_burst"] / 6.0
    if kernel != "3xtf32":
        peak = info["sm_count"] * 128 * 2 * info["sm_clock_khz"] * 1e3 / 1e12
    achieved = flops / avg_ms / 1e9
    cpu_tf, cores, kind, sample, _ = cpu_reference_arm(wl, seconds_target=12.0, cores=1 if wl_name == "c1" else None)
    traffic, traffic_source = k1_traffic_bytes(wl_name, live=not args.no_ncu)
    line = {
        "metric": "SGEMM TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_for(wl_name, 1),
        "details": {"kernel": f"{kernel} (auto dispatch)", "device": info["name"],
                    "accuracy": "3xTF32 with fp32 promotion every 128 k: normwise relerr <= 2.6e-6 vs fp64 on U[0,1) inputs (gate 1e-5)"},
        "e2e": {"value": flops / e2e_ms / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": (M * K + K * N) * 4, "d2h_bytes_per_step": M * N * 4,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "api": "sgemm_cuda(host pointers, pinned)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_source, "algorithmic_bytes": 4 * (M * K + K * N + M * N),
                     "peak_basis": (f"{peaks['source']} bf16 dense burst {peaks['bf16_burst']:.1f} TFLOP/s / 2 (TF32) / 3 (3xTF32)" if kernel == "3xtf32"
                                    else "FP32 FFMA: SMs x 128 lanes x 2 flop x max SM clock"),
                     "kernel_ms_avg": avg_ms, "kernel_ms_min": min_ms,
                     "nominal_frac": achieved / 375.0},
        "cpu_baseline": {"value": cpu_tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
    }
    if sustained is not None:
        line["sustained"] = sustained
    if c5_1gpu is not None:
        line["c5_1gpu"] = c5_1gpu
    if config1 is not None:
        line["config1_host"] = config1
    if shapes is not None:
        line["other_shapes"] = shapes
    print(json.dumps(line), flush=True)


def run_multi(args, wl_name):
    """N > 1: the sharded SGEMM of the C ABI (csrc/shard.cu, sgemm_cuda_shard_*), one process per GPU.  torch.distributed is the
    RENDEZVOUS only (it hands rank 0's 128-byte NCCL id to every rank); communicators, buffers, streams, events, both transports
    and the timing live behind the C ABI."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import ugemm_b200 as u
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # the panel broadcasts only have to keep up with the GEMM of the previous slab, not saturate NVLink: a few CTAs
    # per communicator are enough and fit in the SMs the sharded driver leaves free (csrc/shard.cu)
    os.environ.setdefault("NCCL_MAX_CTAS", "4")
    # stdout carries exactly one JSON line.  NCCL printf()s its version banner to stdout at NCCL_DEBUG=VERSION / WARN (the image
    # sets VERSION), so file descriptor 1 points at stderr for the whole run and the JSON line is written to the saved stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    u.sgemm_cuda_init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [u.Shard.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    wl = WORKLOADS[wl_name]
    M, N, K = wl["M"], wl["N"], wl["K"]
    flops = 2.0 * M * N * K
    want = u.Shard.P2P if args.dist == "p2p" else u.Shard.NCCL
    sh = u.Shard(rank, world, box[0], M, N, K, transport=want)
    plan = sh.p
    if want == u.Shard.P2P and sh.transport != u.Shard.P2P and rank == 0:
        print("peer-pull transport unavailable on this box (no CUDA IPC peer path); every rank fell back to NCCL broadcast", file=sys.stderr, flush=True)
    sh.generate(seed_a=1, seed_b=2)

    def timed(distribute, steps, warmup):
        return sh.allreduce(sh.run(distribute, steps, warmup), "max") / steps      # device time, max over ranks

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = u.launch_count()
    ms = timed(True, args.steps, args.warmup)
    launches = u.launch_count() - l0 - args.warmup * plan["L"]
    clocks = sampler.stop() if sampler else None
    ms_compute = timed(False, max(2, args.steps // 2), 1)
    transport = "p2p" if sh.transport == u.Shard.P2P else "nccl"

    # ---- end to end, pipelined behind the C ABI: owned slabs start in pinned host memory every step, the C block ends there
    sh.download_owned()
    e2e_steps = max(2, min(args.steps, 5))
    e2e_local, h2d, d2h = sh.run_host(e2e_steps, 1)
    e2e_ms = sh.allreduce(e2e_local, "max") / e2e_steps
    e2e_c_keep = sh.host_c().copy()
    floor_ms = sh.allreduce(sh.copy_floor(3), "max") / 3        # the same bytes up and down with nothing else going on
    # bytes per step over all ranks: every slab of A and B goes up exactly once (on its owner), every C block comes down once
    h2d_all, d2h_all = 4 * (M * K + K * N), 4 * M * N
    assert abs(sh.allreduce(float(h2d) / 2 ** 20, "sum") - h2d_all / 2 ** 20) < 1.0 and abs(sh.allreduce(float(d2h) / 2 ** 20, "sum") - d2h_all / 2 ** 20) < 1.0
    e2e_c = e2e_c_keep

    # ---- sampled verification of this rank's C block against fp64 dot products of regenerated windows
    sh.run(True, 1, 0)
    c_ptr, rows, cols, r0, c0 = sh.block()
    rs = np.linspace(0, rows - 1, 6).astype(int)
    cs = np.linspace(0, cols - 1, 48).astype(int)
    full_rows = np.empty((len(rs), cols), np.float32)
    for i, r in enumerate(rs):
        u.backend.lib().ugemm_cuda_memcpy_d2h(full_rows[i].ctypes.data, c_ptr + 4 * int(r) * cols, 4 * cols)
    a_rows = np.stack([u.fill_uniform_host_2d(1, K, 1, (r0 + int(r)) * K, K) for r in rs]).astype(np.float64)
    b_cols = np.stack([u.fill_uniform_host_2d(K, 1, 2, c0 + int(c), N) for c in cs], axis=1).astype(np.float64)
    ref = a_rows @ b_cols
    verr = float(np.linalg.norm(full_rows[:, cs] - ref) / np.linalg.norm(ref))
    e2e_rows = e2e_c.reshape(rows, cols)[rs][:, cs]
    verr_e2e = float(np.linalg.norm(e2e_rows - ref) / np.linalg.norm(ref))
    verr = sh.allreduce(max(verr, verr_e2e), "max")
    if not verr <= 1e-5:
        raise SystemExit(f"sharded result failed verification: sampled relerr {verr:.3e} > 1e-5")

    # ---- the other transport, same problem, same process (north_star names NCCL broadcast; the headline uses whichever --dist asks for)
    other_ms = other_name = None
    if world > 1:
        sh.finish()
        other = u.Shard.NCCL if sh.transport == u.Shard.P2P else u.Shard.P2P
        box = [u.Shard.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sh2 = u.Shard(rank, world, box[0], M, N, K, transport=other)
        if sh2.transport == other:
            sh2.generate(seed_a=1, seed_b=2)
            steps2 = max(2, args.steps // 2)
            other_ms = sh2.allreduce(sh2.run(True, steps2, 2), "max") / steps2
            other_name = "p2p" if other == u.Shard.P2P else "nccl"
        sh2.finish()

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["bf16_burst"] / 6.0 * world
        recv = sum(plan["mloc"] * plan["kw"] * 4 for t in range(plan["L"]) if u.Shard.owners(world, rank, M, N, K, t)[0] != rank) + \
               sum(plan["kw"] * plan["nloc"] * 4 for t in range(plan["L"]) if u.Shard.owners(world, rank, M, N, K, t)[1] != rank)
        details = {"grid": f"{plan['pr']}x{plan['pc']}", "k_slabs": plan["L"], "transport": transport,
                   "timed_region": "owner-rooted panel distribution (%s) + local GEMMs, distribution included, device time, max over ranks"
                                   % ("NCCL broadcast in grid-row / grid-column communicators" if transport == "nccl" else "copy-engine peer pull over NVLink"),
                   "driver": "C ABI (sgemm_cuda_shard_*): NCCL via dlopen, torch.distributed used for the rendezvous of the NCCL id only",
                   "compute_only_tflops": flops / ms_compute / 1e9, "compute_only_ms": ms_compute,
                   "recv_bytes_per_rank": recv, "verified_sampled_relerr_max_over_ranks": verr}
        if other_ms is not None:
            details[f"{other_name}_ms_per_step"] = other_ms
            details[f"{other_name}_tflops"] = flops / other_ms / 1e9
        details[("nccl_broadcast_ms" if transport == "nccl" else "p2p_pull_ms")] = ms
        if other_ms is not None:
            details[("nccl_broadcast_ms" if other_name == "nccl" else "p2p_pull_ms")] = other_ms
        line = {
            "metric": "SGEMM TFLOP/s", "value": flops / ms / 1e9, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_for(wl_name, world),
            "details": details,
            "e2e": {"value": flops / e2e_ms / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "host_link_floor_ms": floor_ms, "frac_of_host_link_floor": floor_ms / e2e_ms,
                    "api": "sgemm_cuda_shard_run_host: owned slabs from pinned host memory, NCCL broadcast, products and the C block's way back pipelined slab by slab; host wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": flops / ms_compute / 1e9, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / ms_compute / 1e9 / peak, "traffic": None,
                         "peak_basis": f"{world} x {peaks['source']} bf16 dense burst {peaks['bf16_burst']:.1f} / 6; achieved = compute-only (panels resident)"},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None] + list(WORKLOADS))
    ap.add_argument("--no-shapes", action="store_true", help="skip the side table of the other BASELINE shapes (N = 1 only)")
    ap.add_argument("--no-ncu", action="store_true", help="do not re-run under ncu for roofline.traffic (use the committed capture)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--dist", default=os.environ.get("UGEMM_BENCH_DIST", "p2p"), choices=["nccl", "p2p"],
                    help="panel transport for --gpus N > 1: NCCL broadcast or copy-engine peer pull")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl_name = args.workload or ("c2" if max(args.gpus, world) == 1 else "c5")
    if args.traffic_child:
        return run_traffic_child(wl_name)
    if args.impl == "reference":
        return run_reference_impl(args, wl_name)
    if world > 1:
        return run_multi(args, wl_name)
    if args.gpus > 1:
        sys.exit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                 "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    run_single(args, wl_name)


if __name__ == "__main__":
    main()
