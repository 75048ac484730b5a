#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2i_pytest.log
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --no-sweep --sustain-seconds 0"
K='hist_kernel|fused_filter|order_|onehot|profile_|kmer_|batched_|tc_|seqmask|resolve|filter_q8|refine|dense_w'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 300 --csv --log-file gpurun_out/r2i_c4_launches.csv \
    python bench.py $FAST > gpurun_out/r2i_c4_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 300 --csv --log-file gpurun_out/r2i_c5_launches.csv \
    python bench.py --workload c5 --n-per-gpu 125000000 $FAST > gpurun_out/r2i_c5_ncu1.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err
tail -2 gpurun_out/r2i_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench_default.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d.get("e2e_api", {}).get("value"), d.get("e2e_api", {}).get("ms_per_call"))
print(d["sustained"], d["clocks"])
for r in d["threshold_sweep_rank0"]:
    print(r)
for k, v in d["other_workloads"].items():
    print(k, {a: b for a, b in v.items() if a in ("value", "ms_per_step", "kernel_ms", "frac", "e2e", "frac_executed")})
print(d["cpu_baseline"])
PY
