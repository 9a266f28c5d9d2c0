#!/bin/bash
# DRAM traffic of K1 at c2 (8192^3) under the L2 eviction-hint settings of the TMA loads (UGEMM_K1_FLAGS bits 7-10):
# ncu, three metrics only, 3 launches each.  Usage (GPU box): bash tools/gpu_traffic.sh  -> gpurun_out/traffic_hints.csv
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/c2_once.py <<'PY'
import sys
sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
n = 8192
dA, dB, dC = u.DeviceBuffer(n * n), u.DeviceBuffer(n * n), u.DeviceBuffer(n * n)
dA.fill_uniform(1); dB.fill_uniform(2)
for _ in range(4):
    u.sgemm_cuda_dev("3xtf32", None, "R", "N", "N", n, n, n, 1.0, dA, n, dB, n, 0.0, dC, n)
u.sync()
PY
: > $OUT/traffic_hints.csv
for F in 1 129 257 385 513 1025 641; do
  echo "# UGEMM_K1_FLAGS=$F" >> $OUT/traffic_hints.csv
  UGEMM_K1_FLAGS=$F timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:k1_3xtf32 -s 1 -c 3 --csv python /tmp/c2_once.py 2>/dev/null | grep -E "k1_3xtf32" | awk -F'","' '{print $(NF-2)","$(NF-1)","$NF}' >> $OUT/traffic_hints.csv
done
cat $OUT/traffic_hints.csv
