"""CPU suite: the C-ABI library loads, exports every symbol include/ugemm_cuda.h declares, and fails loudly
(never falls back) when no GPU is present.  No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ugemm_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:sgemm_cuda|ugemm_cuda|ugemm_fill|im2col_cuda|convolution_cuda|saxpy_cuda|sgemv_cuda|dgemm_cuda)\w*)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_expected_boundary():
    syms = declared_symbols()
    for must in ("sgemm_cuda_init", "sgemm_cuda_finish", "sgemm_cuda", "sgemm_cuda_3xtf32", "sgemm_cuda_simt",
                 "sgemm_cuda_dev", "sgemm_cuda_last_error", "ugemm_fill_uniform_host", "ugemm_fill_uniform_dev"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import ugemm_b200 as u
    L = u.lib()
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/ugemm_cuda.h but not exported"
    from ugemm_b200.backend import EXPORTED_SYMBOLS
    assert sorted(EXPORTED_SYMBOLS) == declared_symbols()


def test_header_compiles_as_plain_c(tmp_path):
    """The boundary must be consumable from the reference's own language (C, gcc)."""
    src = tmp_path / "t.c"
    src.write_text('#include "ugemm_cuda.h"\n'
                   "typedef void (*uut_t)(char, char, char, int, int, int, float, const float*, int, const float*, int, float, float*, int);\n"
                   "int main(void){ uut_t f = sgemm_cuda; uut_t g = sgemm_cuda_3xtf32; uut_t h = sgemm_cuda_simt; return !(f && g && h); }\n")
    subprocess.check_call(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_gpu_means_loud_failure_not_fallback():
    import ugemm_b200 as u
    L = u.lib()
    if L.sgemm_cuda_init(-1, 0) == 0:
        pytest.skip("a GPU is present; the no-GPU path cannot be exercised here")
    L.sgemm_cuda_clear_error()
    A = np.ones(4, np.float32)
    Cm = np.full(4, 7.0, np.float32)
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda("R", "N", "N", 2, 2, 2, 1.0, A, 2, A, 2, 0.0, Cm, 2)
    assert np.array_equal(Cm, np.full(4, 7.0, np.float32)), "C must be untouched when the CUDA path is unavailable"
    with pytest.raises(u.UgemmCudaError):
        u.sgemm_cuda_simt("R", "N", "N", 2, 2, 2, 1.0, A, 2, A, 2, 0.0, Cm, 2)


def test_mgpu_entry_points_fail_loudly_without_init_or_gpu():
    """sgemm_cuda_mgpu before sgemm_cuda_mgpu_init is an error that leaves C untouched (needs no GPU to check);
    without a GPU sgemm_cuda_mgpu_init itself fails with the sticky error set -- no CPU path behind it either."""
    import ugemm_b200 as u
    L = u.lib()
    L.sgemm_cuda_clear_error()
    assert u.sgemm_cuda_mgpu_count() == 0
    A = np.ones(4, np.float32)
    Cm = np.full(4, 7.0, np.float32)
    with pytest.raises(u.UgemmCudaError, match="sgemm_cuda_mgpu_init"):
        u.sgemm_cuda_mgpu("R", "N", "N", 2, 2, 2, 1.0, A, 2, A, 2, 0.0, Cm, 2, 1, 1)
    assert np.array_equal(Cm, np.full(4, 7.0, np.float32))
    assert u.visible_gpus() >= 0
    if u.visible_gpus() == 0:
        with pytest.raises(u.UgemmCudaError):
            u.sgemm_cuda_mgpu_init(1)
        assert u.sgemm_cuda_mgpu_count() == 0
    u.sgemm_cuda_mgpu_finish()      # harmless when nothing was initialised


def test_product_never_references_the_oracle():
    """Nothing under ugemm_b200/ or include/ may import, link or mention oracle/ (checker isolation)."""
    bad = []
    for base in ("ugemm_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle|_ref/|libugemm_ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(ROOT, "ugemm_b200", "libugemm_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ugemm_ref" not in out


def test_k1_schedule_partition_invariants():
    """sgemm_cuda_k1_plan / _plan_item (pure host arithmetic, the function the kernel's roles decode their work with): whole tiles
    appear once, the stream-K tail's chunk ranges tile every tail tile's K extent exactly once in multiples of the promotion
    interval, workspace slots are distinct, and there are never more ranges than CTA pairs (one range per pair)."""
    import random
    import ugemm_b200 as u
    rng = random.Random(7)
    shapes = [(8192, 8192, 8192), (4095, 3001, 2047), (4096, 4096, 4096), (2560, 2560, 2560), (1024, 1024, 1024), (512, 512, 4096),
              (200704, 256, 1152), (1100, 900, 1024), (768, 640, 4096), (129, 257, 33), (128, 128, 32)]
    shapes += [(rng.randint(1, 6000), rng.randint(1, 6000), rng.randint(1, 9000)) for _ in range(150)]
    tails = 0
    try:
        for kc in (4, 2, 8, 0):
            u.set_k1_tuning(kc_blocks=kc)
            for (M, N, K) in shapes:
                for sms in (148, 140, 4, 3):
                    p = u.k1_plan(M, N, K, 1, sms)
                    cg, nkb, kcp = p["cta_group"], p["k_blocks"], p["kc"]
                    assert cg in (1, 2) and p["tile_m"] == 128 * cg and p["tiles_m"] == -(-M // p["tile_m"]) and p["tiles_n"] == -(-N // p["tile_n"])
                    nt = p["tiles_m"] * p["tiles_n"]
                    assert nkb == -(-K // 32) and 1 <= kcp <= nkb
                    assert p["whole_tiles"] + p["tail_tiles"] == nt
                    ranges = p["items"] - p["whole_tiles"]
                    if p["tail_tiles"] == 0:
                        assert p["items"] == nt
                        continue
                    tails += 1
                    assert 0 < ranges <= max(1, sms // cg) and p["whole_tiles"] % max(1, sms // cg) == 0
                    assert u.k1_plan_item(p, 0, 0) == (0, 0, nkb, -1) if p["whole_tiles"] else True
                    cover, slots = {}, set()
                    for item in range(p["whole_tiles"], p["items"]):
                        segs = [u.k1_plan_item(p, item, h) for h in (0, 1)]
                        assert segs[0][2] > segs[0][1], "a range has at least its first segment"
                        for (tile, kb0, kb1, slot) in segs:
                            if kb1 <= kb0:
                                continue
                            assert p["whole_tiles"] <= tile < nt and 0 <= kb0 < kb1 <= nkb and kb0 % kcp == 0 and (kb1 % kcp == 0 or kb1 == nkb)
                            assert 0 <= slot < 2 * ranges and slot not in slots
                            slots.add(slot)
                            cover.setdefault(tile, []).append((kb0, kb1))
                    assert sorted(cover) == list(range(p["whole_tiles"], nt))
                    for tile, spans in cover.items():
                        spans.sort()
                        assert spans[0][0] == 0 and spans[-1][1] == nkb and all(a[1] == b[0] for a, b in zip(spans, spans[1:])), (M, N, K, tile, spans)
    finally:
        u.set_k1_tuning(kc_blocks=4)
    assert tails > 50, tails          # the sweep did exercise the tail
    assert u.last_error() is None
