#!/bin/bash
# C2 (sequence-only scan, computed background): parity of the one-hot paths, the bench line, a device timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -x -q  2>&1 | tail -4 | tee gpurun_out/c2_pytest.log
timeout 300 python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu-baseline --trace gpurun_out/c2_timeline.json > gpurun_out/c2_bench_1.json 2> gpurun_out/c2_bench_1.err
tail -3 gpurun_out/c2_bench_1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c2_bench_1.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
t = json.load(open("gpurun_out/c2_timeline.json"))["events"]
for e in t[-14:]:
    print("%9.1f %7.1f gap %6.1f  %s" % (e["start"], e["dur"], e["gap_before"], e["name"][:60]))
PY
timeout 300 python bench.py --workload c3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/c3_bench_1.json 2> gpurun_out/c3_bench_1.err
python - <<PY
import json
d = json.load(open("gpurun_out/c3_bench_1.json"))
print("c3", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
PY
