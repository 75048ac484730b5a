#!/bin/bash
# one gpurun call: GPU test suite, then the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d.get("e2e_api"))
print(d["sustained"])
for r in d["threshold_sweep_rank0"]:
    print(r)
for k, v in d["other_workloads"].items():
    print(k, {a: b for a, b in v.items() if a in ("value", "ms_per_step", "kernel_ms", "frac", "e2e")})
PY
