#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/cli_ranks_check.py 2> gpurun_out/n${N}_cli.err | tail -1 | tee gpurun_out/n${N}_cli.json
bash tools/gpu_nbench.sh $N
nproc; free -g | head -2
