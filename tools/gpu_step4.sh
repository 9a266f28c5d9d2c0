#!/bin/bash
timeout 900 python -m pytest tests/test_dgemm.py -m gpu -x -q -s 2>&1 | tail -15
rm -f gpurun_out/dgemm.jsonl
timeout 600 python tools/gpu_dgemm.py | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['ta'],d['tb'],d['M'],d['N'],d['K'],'ms %.3f TF avg %.2f best %.2f frac %.2f'%(d['ms_avg'],d['tflops_avg'],d['tflops_best'],d['frac']))"
