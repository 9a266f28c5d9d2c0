#!/bin/bash
# ncu --set full captures for the evidence north_star asks for: K2 FMA-pipe utilisation, K1 on the skinny c4 shape (HBM GB/s)
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import ugemm_b200 as u
mode, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N)
avg,best=u.sgemm_cuda_time_dev(mode, 3, 2, "R","N","N",M,N,K,1.0,dA,K,dB,N,0.0,dC,N)
print(mode, M, N, K, "avg ms", avg, "TF", 2*M*N*K/avg/1e9)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_simt -s 2 -c 1 -f -o $OUT/r1c_k2_c2 python /tmp/one.py simt 8192 8192 8192 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_3xtf32 -s 2 -c 1 -f -o $OUT/r1c_k1_c4 python /tmp/one.py 3xtf32 200704 256 1152 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_simt -s 2 -c 1 -f -o $OUT/r1c_k2_c4 python /tmp/one.py simt 200704 256 1152 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_simt -s 2 -c 1 -f -o $OUT/r1c_k2_skinny python /tmp/one.py simt 200704 16 1152 2>&1 | tail -2
python /tmp/one.py simt 200704 256 1152; python /tmp/one.py simt 200704 16 1152; python /tmp/one.py 3xtf32 200704 256 1152; python /tmp/one.py simt 4095 3001 2047
ls -la $OUT/*.ncu-rep
