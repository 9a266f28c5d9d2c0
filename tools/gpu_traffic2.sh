#!/bin/bash
# DRAM traffic and time of one 8192^3 launch of the TS kernel: serpentine K on (default) / off (bit 19)
OUT=gpurun_out/${1:-r3}_traffic.csv; : > $OUT
cat > /tmp/c2_once.py <<'PY'
import sys; sys.path.insert(0, ".")
import ugemm_b200 as u
u.sgemm_cuda_init()
M = N = K = 8192
dA, dB, dC = u.DeviceBuffer(M * K).fill_uniform(1), u.DeviceBuffer(K * N).fill_uniform(2), u.DeviceBuffer(M * N)
avg, best = u.sgemm_cuda_time_dev("3xtf32", 6, 2, "R", "N", "N", M, N, K, 1.0, dA, K, dB, N, 0.0, dC, N)
print("ms_avg %.4f ms_min %.4f" % (avg, best))
PY
for F in $FL; do
  echo "flags $F" >> $OUT
  UGEMM_K1_FLAGS=$F timeout 60 python /tmp/c2_once.py >> $OUT 2>&1
  UGEMM_K1_FLAGS=$F timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none --print-units base -k regex:k1ts_kernel -s 3 -c 2 --csv python /tmp/c2_once.py 2>/dev/null | grep -E "k1ts" | awk -F'","' '{print $(NF-2)","$(NF-1)","$NF}' >> $OUT
done
cat $OUT
