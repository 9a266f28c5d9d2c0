#!/bin/bash
# one ncu --set full capture of K2 (source-level stall sampling) at 4096^3 NN + narrow-tile timings
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r1f}
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import ugemm_b200 as u
mode, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N)
avg,best=u.sgemm_cuda_time_dev(mode, 3, 2, "R","N","N",M,N,K,1.0,dA,K,dB,N,0.0,dC,N)
print(mode, M, N, K, "avg ms", avg, "TF", 2*M*N*K/avg/1e9)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_simt -s 2 -c 1 -f -o $OUT/${TAG}_k2_4096 python /tmp/one.py simt 4096 4096 4096 2>&1 | tail -2
for n in 8 16 32 64; do python /tmp/one.py simt 200704 $n 1152; done
