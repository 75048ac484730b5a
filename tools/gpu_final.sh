#!/bin/bash
# Round-end evidence run: tests, smoke, default bench (+others), reference arm, ncu of the four headline kernels.
set -u
TAG=${1:-r1g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
for WL in c4 c2 c3 c5; do
  case $WL in c4) K=fused_filter;; c2) K=kmer_scan_kernel;; c3) K=dense_w_kernel;; c5) K=batched_tc_kernel;; esac
  CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e"
  $CMD > gpurun_out/${TAG}_${WL}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_${WL} $CMD > gpurun_out/${TAG}_${WL}_ncu.log 2>&1
  $CMD > gpurun_out/${TAG}_${WL}_plain2.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hist_kernel|fused_filter|order_|refine|onehot|profile_|kmer_|batched_|tc_|dense_w|mask_scan|provisional' -c 400 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu1.log 2>&1
done
# bench lines last, so that profiles/traffic.json (written here from the captures above by tools/ncu_summary.py on the
# build machine) can be picked up by a later run; these lines carry traffic from the previously committed capture
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -2 gpurun_out/${TAG}_bench_default.err
for WL in c2 c3 c5; do timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-others > gpurun_out/${TAG}_bench_$WL.json 2>/dev/null; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-others --no-cpu-baseline --serial-bg > gpurun_out/${TAG}_bench_c4_serial.json 2>/dev/null
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-others --no-cpu-baseline --serial-bg > gpurun_out/${TAG}_bench_c2_serial.json 2>/dev/null
timeout 600 python bench.py --n-per-gpu 1000000000 --steps 5 --warmup 3 --no-others --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_c4_1gnt.json 2>/dev/null
RNASCAN_REF_STEP_SECONDS=6 timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
ls gpurun_out | grep ${TAG} | head -60
