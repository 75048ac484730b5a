#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_hostscan_gpu.py -x -q 2>&1 | tail -4
for F in q4 q8; do
python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline --no-sweep --sustain-seconds 0 --no-api --e2e-form $F > gpurun_out/r2l_bench_$F.json 2> gpurun_out/r2l_bench_$F.err
tail -2 gpurun_out/r2l_bench_$F.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2l_bench_$F.json"))
print(d["value"], d["e2e"])
PY
done
