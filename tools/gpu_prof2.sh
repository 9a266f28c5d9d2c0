#!/bin/bash
export EXPLORE_LOG=prof2.jsonl
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/clocks_prof2.csv &
SMI=$!
echo "=== MMA free-run (flags=96)"; UGEMM_K1_FLAGS=96 timeout 120 python tools/gpu_explore.py time 3xtf32 2 2 0 8192 8192 8192 2>&1 | grep -E "k1prof cta0 |tflops" | tail -2 | cut -c1-420
echo "=== MMA free-run kc=0 (flags=96)"; UGEMM_K1_FLAGS=96 timeout 120 python tools/gpu_explore.py time 3xtf32 2 0 0 8192 8192 8192 2>&1 | grep -E "k1prof cta0 |tflops" | tail -2 | cut -c1-420
echo "=== MMA free-run 1 MMA/kstep (flags=104)"; UGEMM_K1_FLAGS=104 timeout 120 python tools/gpu_explore.py time 3xtf32 2 0 0 8192 8192 8192 2>&1 | grep -E "k1prof cta0 |tflops" | tail -2 | cut -c1-420
echo "=== MMA free-run cg=1 (flags=96)"; UGEMM_K1_FLAGS=96 timeout 120 python tools/gpu_explore.py time 3xtf32 1 0 0 8192 8192 8192 2>&1 | grep -E "k1prof cta0 |tflops" | tail -2 | cut -c1-420
echo "=== cuBLAS peaks"; timeout 300 python tools/gpu_peak.py 2>&1 | tail -1
echo "=== K1 sustained 200 iters"; python - <<'PY'
import sys; sys.path.insert(0,'.')
import ugemm_b200 as u
M=N=K=8192
dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N)
for it in (20, 200, 600):
    avg,best,tot=u.sgemm_cuda_time_dev("3xtf32", it, 3, "R","N","N",M,N,K,1.0,dA,K,dB,N,0.0,dC,N,total=True)
    print(f"iters={it} avg_ms={avg:.3f} min_ms={best:.3f} TF(total)={2*M*N*K*it/tot/1e9:.1f}")
PY
kill $SMI; sort -t, -k2 -n -r gpurun_out/clocks_prof2.csv | head -5; echo; awk -F, '{print $1}' gpurun_out/clocks_prof2.csv | sort | uniq -c | sort -k1 -n -r | head -8
