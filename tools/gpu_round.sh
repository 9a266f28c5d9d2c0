#!/bin/bash
# One GPU-box pass: parity tests, smoke, measured cuBLAS peaks, bench (both arms), ncu launch list + full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [skip_tests] [skip_ncu]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/${TAG}_gpu.txt
if [ -z "$2" ] || [ "$2" = "0" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
  grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -5
  timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
fi
timeout 300 python tools/gpu_peak.py > $OUT/${TAG}_cublas_peaks.jsonl 2> $OUT/${TAG}_cublas_peaks.err; echo "peaks rc=$?"; cat $OUT/${TAG}_cublas_peaks.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "bench ref rc=$?"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; head -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [ -z "$3" ] || [ "$3" = "0" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-ncu > $OUT/${TAG}_ncu_launches_stdout.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k1ts_kernel|k1_3xtf32" -s 4 -c 1 -f -o $OUT/${TAG}_k1_c2 python bench.py --steps 5 --warmup 3 --no-shapes --no-ncu > $OUT/${TAG}_ncu_full_stdout.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT | tail -20
