"""The C harnesses above the C ABI (harness/*.c: check_sgemm.c / sgemm_test.c extended to the CUDA backend).

CPU part: the programs build with gcc against include/ugemm_cuda.h, and check_sgemm_cuda's host-only pass (`mode=cpu`, BASELINE
config 1: the reference's own uut rows on this host, no GPU touched) runs green over the reference's 11 stacked instances.
GPU part (-m gpu): every harness binary runs on the box -- config 1 with 11 instances, config 3 with a transpose (padded and odd
leading dimensions), the sgemm_test.c macro harness, the DGEMM harness, the single-process sharded harness and the one-process-per-GPU
sharded harness -- exit status 0 and no failed gate line."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "harness")
BINARIES = ("check_sgemm_cuda", "sgemm_test_cuda", "check_dgemm_cuda", "sgemm_mgpu_cuda", "sgemm_shard_cuda")


@pytest.fixture(scope="module")
def harness():
    from ugemm_b200 import build as b
    b.build()                                     # the product library the programs link against
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], env=env, stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HARNESS, "all"], env=env, stdout=subprocess.DEVNULL)
    for name in BINARIES:
        assert os.access(os.path.join(HARNESS, name), os.X_OK), name
    return HARNESS


def run(harness, name, *args, timeout=600):
    res = subprocess.run([os.path.join(harness, name), *map(str, args)], cwd=harness, capture_output=True, text=True, timeout=timeout)
    print(f"$ {name} {' '.join(map(str, args))}  -> rc {res.returncode}")
    # the cmp_results lines of the reference format are many; keep the verdict lines
    print("\n".join(ln for ln in res.stdout.splitlines() if "e-0" not in ln or "relerr" in ln), res.stderr[-2000:])
    return res


def green(res):
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "FAIL (>" not in res.stdout            # (the reference-format lines print the reference's own, broken, "FAIL !!!" rule)
    assert "padding written" not in res.stdout


def test_host_only_config1_pass(harness):
    """`mode=cpu`: the reference rows of check_sgemm.c:247-250 over 11 stacked instances, compared with the naive ground truth;
    the CUDA rows are reported not run (there is no CPU fallback behind them)."""
    res = run(harness, "check_sgemm_cuda", "mode=cpu", "M=256", "N=250", "K=300", "inst=11", "ldc=256")
    green(res)
    out = res.stdout
    assert "not run (mode=cpu" in out
    for row in ("sgemm_c (gemm_cpu.h:284)", "sgemm_avx 1 core (as shipped)", "sgemm_sse (sgemm_sse.h:365)", "sgemm_avx row slabs, all cores"):
        assert f"{row:<28s} worst normwise relerr over 11 instances" in out, row
    assert out.count(" ok\n") >= 4


def test_cuda_rows_fail_loudly_without_a_gpu(harness):
    import ugemm_b200 as u
    if u.visible_gpus() > 0:
        pytest.skip("a GPU is visible: the loud-failure branch is exercised on the CPU-only builder")
    res = run(harness, "check_sgemm_cuda", "M=64", "N=64", "K=64")
    assert res.returncode == 1 and "sgemm_cuda_init" in res.stderr


@pytest.mark.gpu
def test_check_sgemm_cuda_config1_eleven_instances(harness):
    """BASELINE config 1 shape on the GPU rows: 1024^3 NN, the reference's 11 stacked instances walked per call and in one
    sgemm_cuda_batched launch, each instance against sgemm_avx."""
    res = run(harness, "check_sgemm_cuda", "M=1024", "N=1024", "K=1024", "inst=11", "iters=2")
    green(res)
    for row in ("sgemm_cuda_3xtf32", "sgemm_cuda_simt", "sgemm_cuda (auto)", "sgemm_cuda_batched"):
        assert f"{row:<28s} worst normwise relerr over 11 instances" in res.stdout, row


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ("ta=T", "lda=4096", "ldb=3004", "ldc=3004"),                 # TN, padded to multiples of 4 -> K1
    ("tb=T", "lda=2052", "ldb=2050", "ldc=3008"),                 # NT, ldb not a multiple of 4: forced K1 reports, auto repacks
    ("ta=T", "tb=T", "lda=4100", "ldb=2050", "ldc=3008"),         # TT, both odd
    ("major=C", "ta=T", "lda=2048", "ldb=2048", "ldc=4096"),      # column-major
])
def test_check_sgemm_cuda_config3(harness, args):
    res = run(harness, "check_sgemm_cuda", "M=4095", "N=3001", "K=2047", "alpha=1.5", "beta=0.5", "iters=2", *args)
    green(res)
    assert "sgemm_cuda (auto)" in res.stdout and "worst normwise relerr" in res.stdout


@pytest.mark.gpu
def test_sgemm_test_cuda(harness):
    res = run(harness, "sgemm_test_cuda")
    green(res)
    assert "known answer: ok" in res.stdout
    for case in ("RNN", "RNT", "RTN"):
        assert f"{case} normwise relerr" in res.stdout


@pytest.mark.gpu
def test_check_dgemm_cuda(harness):
    green(run(harness, "check_dgemm_cuda", "M=512", "N=384", "K=640"))


@pytest.mark.gpu
def test_sgemm_mgpu_cuda(harness):
    """Config 5's C host program on every GPU of the box at a size that keeps the test short (the full 32768^3 parity check is
    tests/test_parity_gpu.py::test_config5_32768_sampled)."""
    res = run(harness, "sgemm_mgpu_cuda", 0, 8192, 8192, 8192, 1)
    green(res)
    assert "sampled relerr" in res.stdout and " ok" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("transport", [1, 0])
def test_sgemm_shard_cuda(harness, transport):
    """Config 5's C host program with one process per GPU (fork, NCCL id through a shared page, sgemm_cuda_shard_*) on every GPU
    of the box, both transports, at a size that keeps the test short."""
    res = run(harness, "sgemm_shard_cuda", 0, 8192, 8192, 8192, 2, transport)
    green(res)
    assert "sampled relerr" in res.stdout and " ok" in res.stdout and "FAIL" not in res.stdout
