#!/bin/bash
# last check of the round on one GPU: the GPU test suite, smoke(), a short default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/last_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/last_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-others > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
tail -2 gpurun_out/last_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/last_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["form"], d.get("e2e_api", {}).get("value"))
print(d["cpu_baseline"])
PY
