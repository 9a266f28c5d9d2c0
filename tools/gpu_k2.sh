#!/bin/bash
# K2 check: parity subset + timings of the shapes K2 is responsible for
export EXPLORE_LOG=${1:-k2c.jsonl}
timeout 900 python -m pytest tests -m gpu -x -q -k "k2 or simt or K2 or sweep or dispatch or golden or c3 or conv" 2>&1 | tail -3
for s in "8192 8192 8192 N N" "4096 4096 4096 N N" "4096 4096 4096 T N" "4096 4096 4096 N T" "4096 4096 4096 T T" "1024 1024 1024 N N" \
         "200704 256 1152 N N" "200704 64 1152 N N" "200704 32 1152 N N" "200704 16 1152 N N" "200704 8 1152 N N" "4095 3001 2047 N T"; do
  set -- $s
  timeout 120 python tools/gpu_explore.py time simt 2 0 0 $1 $2 $3 $4 $5 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['M'],d['N'],d['K'],d['ta'],d['tb'],'ms %.3f  TF avg %.1f best %.1f'%(d['ms_avg'],d['tflops_avg'],d['tflops_best']))"
done
