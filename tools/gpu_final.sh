#!/bin/bash
# Final pass of a round without the reference arm and the cuBLAS peaks (unchanged since r3z): tests, smoke, bench, ncu launch list, full capture.
TAG=${1:-r4z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 100 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "nominal", round(d["roofline"]["nominal_frac"], 3), "traffic", d["roofline"]["traffic"])
print("sustained", round(d["sustained"]["tflops"], 1), "c5_1gpu", round(d["c5_1gpu"]["ms_avg"], 1), "c1 gpu ms", d["config1_host"]["gpu_ms_avg_l2_cold"])
for s in d["other_shapes"]:
    print("  ", s["shape"][:64], "| ms", round(s["ms_avg"], 4), "| TF", round(s["tflops"], 1), "| frac", round(s.get("roofline_frac", 0), 3))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-ncu --no-shapes > $OUT/${TAG}_ncu_launches_stdout.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k1ts_kernel" -s 4 -c 1 -f -o $OUT/${TAG}_k1_c2 python bench.py --steps 5 --warmup 3 --no-shapes --no-ncu > $OUT/${TAG}_ncu_full_stdout.log 2>&1; echo "ncu full rc=$?"
