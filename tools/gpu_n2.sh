#!/bin/bash
# N ranks on one box: sharded CLI against the golden files, overlapped background pipelines under a real all-reduce, bench
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/cli_ranks_check.py 2> gpurun_out/n${N}_cli.err | tail -1 | tee gpurun_out/n${N}_cli.json
timeout 300 $TR tools/ranks_check_bg.py 2> gpurun_out/n${N}_bg.err | tail -3 | tee gpurun_out/n${N}_bg.log
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
tail -2 gpurun_out/n${N}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/n${N}_bench.json"))
print(d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["link_GBps"], d["config"]["symbols_per_gpu"])
print(d.get("sustained"))
PY
