#!/bin/bash
# One GPU-box session: parity tests, smoke, bench lines, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; tail -3 gpurun_out/${TAG}_bench_c4.err; cat gpurun_out/${TAG}_bench_c4.json
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -3 gpurun_out/${TAG}_bench_c2.err; cat gpurun_out/${TAG}_bench_c2.json
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; tail -3 gpurun_out/${TAG}_bench_c3.err; cat gpurun_out/${TAG}_bench_c3.json
RNASCAN_REF_STEP_SECONDS=4 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -3 gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
if [ "${2:-ncu}" = "ncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hist_kernel|fused_filter|order_|onehot|profile_' -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_filter -s 3 -c 2 -f -o gpurun_out/${TAG}_fused $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/
fi
