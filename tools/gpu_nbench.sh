#!/bin/bash
# bench.py under torchrun on N GPUs of one box (strong scaling: the 1 Gnt of config 4 split N ways)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
tail -2 gpurun_out/n${N}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/n${N}_bench.json"))
print(d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["link_GBps"], d["config"]["symbols_per_gpu"])
print(d.get("sustained"))
print(d.get("e2e_f32"))
PY
