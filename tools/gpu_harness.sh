#!/bin/bash
# C harnesses on the GPU box (the reference's own host language): check_sgemm_cuda at c1 / c3, sgemm_test_cuda, check_dgemm_cuda
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/${1:-r1f}_harness.log; : > $LOG
cd harness
for args in "M=1024 N=1024 K=1024" "M=4095 N=3001 K=2047 ta=T tb=N alpha=1.5 beta=0.5 lda=4096 ldb=3004 ldc=3004" \
            "M=4095 N=3001 K=2047 ta=N tb=T alpha=1.5 beta=0.5 lda=2052 ldb=2050 ldc=3008" "M=8192 N=8192 K=8192 check=1 iters=2"; do
  timeout 600 ./check_sgemm_cuda $args >> ../$LOG 2>&1; echo "rc=$?" >> ../$LOG
done
timeout 300 ./sgemm_test_cuda >> ../$LOG 2>&1; echo "rc=$?" >> ../$LOG
for args in "M=1024 N=1024 K=1024" "M=2048 N=2048 K=2048 alpha=1.5 beta=0.5 lda=2050 ldb=2052 ldc=2054" "M=128 N=361 K=1152"; do
  timeout 600 ./check_dgemm_cuda $args >> ../$LOG 2>&1; echo "rc=$?" >> ../$LOG
done
cd ..; grep -E "rc=|PASSED|FAILED|relerr|TFLOP" $LOG | tail -40
