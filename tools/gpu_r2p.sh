#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "fused" 2>&1 | tail -5
for M in 0 1; do
  RS_BENCH_COUNT_IN_KERNEL=$M python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-sweep --no-e2e > gpurun_out/r2p_bench_cik$M.json 2> gpurun_out/r2p_bench_cik$M.err
  tail -1 gpurun_out/r2p_bench_cik$M.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2p_bench_cik$M.json"))
print("count_in_kernel=$M", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["sustained"]["value"], d["hits_rank0"], d["gpu_launches"])
PY
done
