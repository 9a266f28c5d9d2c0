#!/bin/bash
# Compile k2_simt.cu to a cubin and print, per kernel instantiation: registers, spills, and the instruction mix of the
# hottest loop (the backward-branch region with the most FFMAs).  CPU-only (nvcc cross-compiles).
set -e
OUT=${1:-/tmp/k2}
mkdir -p $OUT
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -cubin \
  -o $OUT/k2.cubin "$(dirname "$0")/../ugemm_b200/csrc/k2_simt.cu" 2> $OUT/ptxas.log
python3 - "$OUT" <<'PY'
import re, subprocess, sys
out = sys.argv[1]
log = open(out + "/ptxas.log").read()
info = {}
for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers.*?(\d+) bytes smem", log):
    info[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)), int(m.group(6)))
sass = subprocess.run(["cuobjdump", "-sass", out + "/k2.cubin"], capture_output=True, text=True).stdout
for fn in re.split(r"\n\s*Function : ", sass)[1:]:
    name = fn.split("\n", 1)[0].strip()
    if "k2_simt_kernel" not in name:
        continue
    tag = re.search(r"k2_simt_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb([01])ELb([01])", name)
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", fn)
    addr = {int(a, 16): i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < int(a, 16) and int(m.group(1), 16) in addr:
            body = ins[addr[int(m.group(1), 16)]:i + 1]
            nf = sum(1 for _, x in body if "FFMA" in x)
            if best is None or nf > best[0]:
                best = (nf, len(body), sum(1 for _, x in body if "LDS" in x), sum(1 for _, x in body if "LDG" in x),
                        sum(1 for _, x in body if "STS" in x), sum(1 for _, x in body if "LDL" in x or "STL" in x))
    r = info.get(name, ("?",) * 4)
    print("%sx%s %sx%s AK%s BK%s: regs %s spill st/ld %s/%s smem %s | hot loop: %s instr, FFMA %s (%.1f%%), LDS %s, LDG %s, STS %s, local %s" % (
        *tag.groups(), r[0], r[1], r[2], r[3], best[1], best[0], 100.0 * best[0] / best[1], best[2], best[3], best[4], best[5]))
PY
