#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_level12.py -m gpu -x -q 2>&1 | tail -5
rm -f $OUT/l12.jsonl
timeout 300 python tools/gpu_l12.py | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], d.get('trans',''), d.get('M',''), d.get('N',d.get('n')), 'ms %.4f GB/s %.0f frac %.2f'%(d['ms'],d['gbs'],d['frac_of_hbm_peak']))"
( for tool in memcheck racecheck synccheck; do echo "=== $tool"; timeout 500 compute-sanitizer --tool $tool python tools/sanitize_cases.py 2>&1 | grep -v "^=========     \|^========= $" | tail -40; done ) > $OUT/r1f_sanitizer.log 2>&1
grep -E "===|ERROR SUMMARY|RACECHECK SUMMARY|hazards|Error" $OUT/r1f_sanitizer.log | sort | uniq -c | head -20
