#!/usr/bin/env python
"""Exploration driver for a GPU box: each experiment runs in its own subprocess (a kernel trap must not
take the others down) under a timeout, and appends JSON lines to gpurun_out/explore.jsonl.

    python tools/gpu_explore.py all            # everything
    python tools/gpu_explore.py probe          # one experiment in-process

Ground truth here is numpy float64 (fast); the parity tests proper (tests/, -m gpu) use the oracle.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = os.path.join(OUT, os.environ.get("EXPLORE_LOG", "explore.jsonl"))


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(LOG, "a") as f:
        f.write(line + "\n")


def f32(bits):
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def hexs(x):
    return ["%08x" % v for v in np.asarray(x, dtype=np.float32).view(np.uint32).ravel()]


def op_matrices(ta, tb, M, N, K, rng, lo, hi, lda_pad=0, ldb_pad=0):
    """Row-major stored A, B (with optional padding) and the logical op(A) (M x K), op(B) (K x N)."""
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    A = rng.uniform(lo, hi, size=(ar, ac + lda_pad)).astype(np.float32)
    B = rng.uniform(lo, hi, size=(br, bc + ldb_pad)).astype(np.float32)
    opA = A[:, :ac] if ta == "N" else A[:, :ac].T
    opB = B[:, :bc] if tb == "N" else B[:, :bc].T
    return A, B, opA, opB


def relerr(x, ref):
    x = x.astype(np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300))


# ------------------------------------------------------------------------------------------------------
def exp_probe():
    import ugemm_b200 as u
    info = u.device_info()
    emit(exp="device", **info)
    rng = np.random.default_rng(0)
    # (1) operand rounding: random fp32 mantissas, one k-step; compare with truncated / RNA / exact models
    A = np.zeros((128, 8), np.float32)
    B = np.zeros((16, 8), np.float32)
    A[:, 0] = rng.uniform(1, 2, 128).astype(np.float32)
    B[:, 0] = rng.uniform(1, 2, 16).astype(np.float32)
    D = u.probe_tf32(A, B, 1)

    def trunc(x):
        return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    def rna(x):
        b = x.view(np.uint32).astype(np.uint64) + 0x1000
        return (b & 0xFFFFE000).astype(np.uint32).view(np.float32)

    a0, b0 = A[:, 0].copy(), B[:, 0].copy()
    models = {"trunc": np.outer(trunc(a0).astype(np.float64), trunc(b0).astype(np.float64)),
              "rna": np.outer(rna(a0).astype(np.float64), rna(b0).astype(np.float64)),
              "exact": np.outer(a0.astype(np.float64), b0.astype(np.float64))}
    res = {k: float(np.abs(D - v).max()) for k, v in models.items()}
    res_rn = {k + "_rn32": bool(np.array_equal(D, v.astype(np.float32))) for k, v in models.items()}
    emit(exp="probe_operand_rounding", max_abs_diff=res, **res_rn, sample=hexs(D[0, :4]))

    # (2) accumulator rounding between MMA instructions (2 k-steps): D = 1*1 then += a*b
    cases = [("plus_1.5ulp_half", 1.5, 2.0 ** -24), ("plus_half_ulp", 1.0, 2.0 ** -24), ("plus_ulp", 1.0, 2.0 ** -23),
             ("minus_quarter_ulp_below1", -1.0, 2.0 ** -26), ("minus_half_ulp_below1", -1.0, 2.0 ** -25),
             ("plus_0.75ulp", 1.5, 2.0 ** -24 * 1.0), ("plus_0.25ulp", 1.0, 2.0 ** -25)]
    A = np.zeros((128, 16), np.float32)
    B = np.zeros((16, 16), np.float32)
    A[:, 0] = 1.0
    B[:, 0] = 1.0
    for i, (_, a, b) in enumerate(cases):
        A[i, 8] = a
    # each row i uses column i of B for its own addend: D[i, i]
    for i, (_, a, b) in enumerate(cases):
        B[i, 8] = b
    D = u.probe_tf32(A, B, 2)
    out = {}
    for i, (name, a, b) in enumerate(cases):
        exact = 1.0 + a * b
        out[name] = {"got": hexs(D[i, i])[0], "rn": hexs(np.float32(exact))[0],
                     "exact": repr(exact)}
    emit(exp="probe_accumulate_rounding_between_instr", cases=out)

    # (3) summation inside one instruction: 1 + 7 * 2^-24 in a single k-step
    A = np.zeros((128, 8), np.float32)
    B = np.zeros((16, 8), np.float32)
    A[:, :] = 1.0
    B[0, 0] = 1.0
    B[0, 1:] = 2.0 ** -24          # 1 + 7*2^-24
    B[1, 0] = 1.0
    B[1, 1:4] = 2.0 ** -24         # 1 + 3*2^-24 -> RN 1+2*2^-24.. (1.5 ulp)
    B[2, 0] = 1.0
    B[2, 1] = 2.0 ** -24           # tie
    B[3, 0] = 1.0
    B[3, 1:] = 2.0 ** -26          # 1 + 7*2^-26 = 1 + 0.875 * 2^-23... (< ulp, > half)
    D = u.probe_tf32(A, B, 1)
    emit(exp="probe_sum_inside_instr", got=hexs(D[0, :4]),
         exact=[repr(1 + 7 * 2.0 ** -24), repr(1 + 3 * 2.0 ** -24), repr(1 + 2.0 ** -24), repr(1 + 7 * 2.0 ** -26)],
         rn=hexs(np.array([1 + 7 * 2.0 ** -24, 1 + 3 * 2.0 ** -24, 1 + 2.0 ** -24, 1 + 7 * 2.0 ** -26], np.float64).astype(np.float32)))

    # (4) chain of 4 k-steps on random data vs fp64 of truncated operands
    A = rng.uniform(0, 1, (128, 32)).astype(np.float32)
    B = rng.uniform(0, 1, (16, 32)).astype(np.float32)
    D = u.probe_tf32(A, B, 4)
    ref_t = trunc(A.copy()).astype(np.float64) @ trunc(B.copy()).astype(np.float64).T
    ref_e = A.astype(np.float64) @ B.astype(np.float64).T
    emit(exp="probe_chain4", relerr_vs_trunc_model=relerr(D, ref_t), relerr_vs_exact=relerr(D, ref_e))


def run_gemm(u, mode, ta, tb, M, N, K, alpha=1.0, beta=0.0, lo=0.0, hi=1.0, pads=(0, 0, 0), seed=0, host_path=True):
    rng = np.random.default_rng(seed)
    A, B, opA, opB = op_matrices(ta, tb, M, N, K, rng, lo, hi, pads[0], pads[1])
    ldc = N + pads[2]
    Cm = rng.uniform(lo, hi, size=(M, ldc)).astype(np.float32)
    C0 = Cm.copy()
    fn = {"auto": u.sgemm_cuda, "3xtf32": u.sgemm_cuda_3xtf32, "simt": u.sgemm_cuda_simt}[mode]
    fn("R", ta, tb, M, N, K, alpha, A.ravel(), A.shape[1], B.ravel(), B.shape[1], beta, Cm.ravel(), ldc)
    ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0[:, :N].astype(np.float64)
    pad_ok = bool(np.array_equal(Cm[:, N:], C0[:, N:]))
    return relerr(Cm[:, :N], ref), pad_ok, Cm, ref


def exp_k2():
    import ugemm_b200 as u
    shapes = [(3, 3, 2), (1, 1, 1), (7, 5, 3), (64, 64, 16), (129, 97, 131), (255, 257, 33), (300, 200, 100),
              (1023, 1000, 1023), (2048, 2048, 512)]
    for (M, N, K) in shapes:
        for ta in "NT":
            for tb in "NT":
                for (alpha, beta, pads) in [(1.0, 0.0, (0, 0, 0)), (1.5, 0.5, (5, 3, 7)), (1.5, 0.5, (4, 4, 4))]:
                    try:
                        e, pad_ok, _, _ = run_gemm(u, "simt", ta, tb, M, N, K, alpha, beta, pads=pads)
                        emit(exp="k2", ta=ta, tb=tb, M=M, N=N, K=K, alpha=alpha, beta=beta, pads=pads, relerr=e, pad_ok=pad_ok,
                             ok=bool(e < 1e-5 and pad_ok))
                    except Exception as ex:
                        emit(exp="k2", ta=ta, tb=tb, M=M, N=N, K=K, error=str(ex))
                        return


def exp_k1(cg, ta, tb):
    import ugemm_b200 as u
    cg = int(cg)
    u.set_k1_tuning(0, 0, cg)
    shapes = [(128 * cg, 128 * cg, 32), (128 * cg, 128 * cg, 64), (256, 256, 256), (512, 768, 320), (300, 200, 100), (1023, 1000, 1023)]
    for (M, N, K) in shapes:
        for (alpha, beta, pads) in [(1.0, 0.0, (0, 0, 0)), (1.5, 0.5, (4, 8, 4))]:
            try:
                e, pad_ok, Cm, ref = run_gemm(u, "3xtf32", ta, tb, M, N, K, alpha, beta, pads=pads)
                extra = {}
                if not (e < 1e-5):
                    d = np.abs(Cm[:, :N].astype(np.float64) - ref)
                    bad = np.argwhere(d > 1e-3 * np.abs(ref).max())
                    extra = {"n_bad": int(len(bad)), "first_bad": bad[:6].tolist(),
                             "bad_rows": sorted(set((bad[:, 0] // 32).tolist()))[:16], "bad_cols": sorted(set((bad[:, 1] // 32).tolist()))[:16],
                             "sample_got": Cm[:2, :4].tolist(), "sample_ref": ref[:2, :4].tolist()}
                emit(exp="k1", cg=cg, ta=ta, tb=tb, M=M, N=N, K=K, alpha=alpha, beta=beta, pads=pads, relerr=e, pad_ok=pad_ok,
                     ok=bool(e < 1e-5 and pad_ok), **extra)
            except Exception as ex:
                emit(exp="k1", cg=cg, ta=ta, tb=tb, M=M, N=N, K=K, error=str(ex))
                return


def exp_k1_accuracy(cg):
    import ugemm_b200 as u
    cg = int(cg)
    M = N = 512
    for K in (1024, 8192, 32768):
        for (lo, hi) in ((0.0, 1.0), (-0.5, 0.5)):
            rows = {}
            for split in (0, 1):
                for kc in (0, 1, 2, 4, 8, 16, 64):
                    u.set_k1_tuning(kc, split, cg)
                    e, _, _, _ = run_gemm(u, "3xtf32", "N", "N", M, N, K, lo=lo, hi=hi, seed=1)
                    rows[f"split{split}_kc{kc}"] = e
            e2, _, _, _ = run_gemm(u, "simt", "N", "N", M, N, K, lo=lo, hi=hi, seed=1)
            rows["simt"] = e2
            emit(exp="k1_accuracy", cg=cg, M=M, N=N, K=K, lo=lo, hi=hi, relerr=rows)


def exp_time(mode, cg, kc, split, M, N, K, ta="N", tb="N"):
    import ugemm_b200 as u
    cg, kc, split, M, N, K = int(cg), int(kc), int(split), int(M), int(N), int(K)
    u.set_k1_tuning(kc, split, cg)
    ar, ac = (M, K) if ta == "N" else (K, M)
    br, bc = (K, N) if tb == "N" else (N, K)
    dA = u.DeviceBuffer(ar * ac).fill_uniform(1)
    dB = u.DeviceBuffer(br * bc).fill_uniform(2)
    dC = u.DeviceBuffer(M * N).fill_uniform(3)
    iters = 10 if mode != "simt" else 3
    avg, best = u.sgemm_cuda_time_dev(mode, iters, 2, "R", ta, tb, M, N, K, 1.0, dA, ac, dB, bc, 0.0, dC, N)
    flops = 2.0 * M * N * K
    emit(exp="time", mode=mode, cg=cg, kc=kc, split=split, M=M, N=N, K=K, ta=ta, tb=tb, ms_avg=avg, ms_min=best,
         tflops_avg=flops / avg / 1e9, tflops_best=flops / best / 1e9)


EXPERIMENTS = {"probe": exp_probe, "k2": exp_k2, "k1": exp_k1, "k1_accuracy": exp_k1_accuracy, "time": exp_time}


def run_all():
    plan = [["probe"], ["k2"]]
    for cg in (1, 2):
        for ta in "NT":
            for tb in "NT":
                plan.append(["k1", str(cg), ta, tb])
    plan += [["k1_accuracy", "1"], ["k1_accuracy", "2"]]
    for cg in (1, 2):
        for kc in (0, 2, 4):
            plan.append(["time", "3xtf32", str(cg), str(kc), "0", "8192", "8192", "8192"])
    plan.append(["time", "3xtf32", "2", "0", "1", "8192", "8192", "8192"])
    plan.append(["time", "3xtf32", "2", "0", "0", "4096", "4096", "4096"])
    plan.append(["time", "3xtf32", "2", "0", "0", "200704", "256", "1152"])
    plan.append(["time", "3xtf32", "2", "0", "0", "8192", "8192", "8192", "T", "N"])
    plan.append(["time", "3xtf32", "2", "0", "0", "8192", "8192", "8192", "N", "T"])
    plan.append(["time", "simt", "2", "0", "0", "8192", "8192", "8192"])
    plan.append(["time", "simt", "2", "0", "0", "4096", "4096", "4096"])
    for args in plan:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + args, timeout=240, capture_output=True, text=True)
            rc, tail = r.returncode, (r.stderr or "")[-600:]
        except subprocess.TimeoutExpired:
            rc, tail = -999, "TIMEOUT"
        emit(exp="_done", args=args, rc=rc, secs=round(time.time() - t0, 1), stderr_tail=tail if rc else "")


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "all":
        run_all()
    else:
        EXPERIMENTS[sys.argv[1]](*sys.argv[2:])
