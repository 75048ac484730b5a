#!/bin/bash
# run-to-run spread of the 20-step timed region at N GPUs (quick bench lines, main workload only)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2 3 4; do
  timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-others --no-sweep --no-cpu-baseline --no-api > gpurun_out/n${N}_var_$i.json 2> gpurun_out/n${N}_var_$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/n${N}_var_$i.json"))
print($i, d["value"], d["ms_per_step"], d["sustained"]["value"], d["sustained"]["ms_per_step"])
PY
done
