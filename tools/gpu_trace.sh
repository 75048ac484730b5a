#!/bin/bash
# device timelines of a few steps of each workload (bench.py --trace; CUPTI through torch.profiler)
mkdir -p gpurun_out
for wl in c4 c5 c3; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-others --no-e2e --no-sweep --no-api --no-cpu-baseline --sustain-seconds 0 --trace gpurun_out/${wl}_timeline.json > gpurun_out/${wl}_trace_bench.json 2> gpurun_out/${wl}_trace.err
  python - <<PY
import json
t = json.load(open("gpurun_out/${wl}_timeline.json"))["events"]
print("== $wl: %d events" % len(t))
k = max(0, len(t) - (14 if "$wl" != "c5" else 40))
for e in t[k:]:
    print("%10.1f %8.1f gap %7.1f  %s" % (e["start"], e["dur"], e["gap_before"], e["name"][:70]))
PY
done
