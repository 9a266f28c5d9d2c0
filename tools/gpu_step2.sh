#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import ugemm_b200 as u
mode, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N)
avg,best=u.sgemm_cuda_time_dev(mode, 3, 2, "R","N","N",M,N,K,1.0,dA,K,dB,N,0.0,dC,N)
print(mode, M, N, K, "avg ms", avg, "TF", 2*M*N*K/avg/1e9)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_simt -s 2 -c 1 -f -o $OUT/r1f_k2_c2 python /tmp/one.py simt 8192 8192 8192 2>&1 | tail -2
rm -f $OUT/l12.jsonl
timeout 600 python tools/gpu_l12.py
