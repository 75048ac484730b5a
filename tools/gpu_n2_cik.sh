#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for M in 0 1; do
  RS_BENCH_COUNT_IN_KERNEL=$M timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-sweep --no-e2e > gpurun_out/n2_cik$M.json 2> gpurun_out/n2_cik$M.err
  python - <<PY
import json
d = json.load(open("gpurun_out/n2_cik$M.json"))
print("N=2 count_in_kernel=$M", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["sustained"]["value"])
PY
done
