#!/bin/bash
# beta != 0: old C through the TMA unit into the staging box, added between promotions (default) vs registers preloaded with
# global loads (bit 22); per-role counters of 1-CTA tiles at 1024^3.  usage: bash tools/gpu_beta_ab2.sh <tag>
TAG=${1:-beta2}
OUT=gpurun_out/${TAG}_beta_ab.jsonl; : > $OUT
for F in ${FLAGSETS:-1 4194305 1 4194305}; do
  UGEMM_K1_FLAGS=$F timeout 60 python tools/gpu_beta_cases.py 2>&1 | tail -1 >> $OUT
done
cat $OUT
cat > /tmp/small.py <<'PY'
import sys; sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
for (M, N, K, cg) in ((1024, 1024, 1024, 0), (1024, 1024, 1024, 2), (2560, 2560, 2560, 1)):
    u.set_k1_tuning(cta_group=cg)
    dA, dB, dC = u.DeviceBuffer(M * K).fill_uniform(1), u.DeviceBuffer(K * N).fill_uniform(2), u.DeviceBuffer(M * N)
    avg, best = u.sgemm_cuda_time_dev("3xtf32", 3, 1, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
    print(f"{M}x{N}x{K} cg={cg}: avg {avg*1e3:.1f} us", file=sys.stderr, flush=True)
PY
[ -n "$NOPROF" ] || UGEMM_K1_FLAGS=33 UGEMM_K1_ABLATION=1 timeout 60 python /tmp/small.py 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${TAG}_small_prof.log
[ -n "$NOPROF" ] || tail -12 gpurun_out/${TAG}_small_prof.log | cut -c1-400
