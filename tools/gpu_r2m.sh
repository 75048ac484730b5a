#!/bin/bash
# final round-2 check on one GPU: test suite, smoke, the driver's bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2m_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2m_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err
tail -2 gpurun_out/r2m_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench_default.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["form"], d.get("e2e_api", {}).get("value"))
print(d["sustained"], d["clocks"])
for k, v in d["other_workloads"].items():
    print(k, {a: b for a, b in v.items() if a in ("value", "ms_per_step", "kernel_ms", "frac", "e2e", "frac_executed")})
PY
timeout 300 python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu-baseline --trace gpurun_out/c2_timeline.json > gpurun_out/c2_bench_1.json 2> gpurun_out/c2_bench_1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c2_bench_1.json"))
print("c2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
t = json.load(open("gpurun_out/c2_timeline.json"))["events"]
for e in t[-7:]:
    print("%9.1f %7.1f gap %6.1f  %s" % (e["start"], e["dur"], e["gap_before"], e["name"][:60]))
PY
