#!/bin/bash
# ncu captures of the three headline kernels (one gpurun call; each command first runs plain).
set -u
TAG=${1:-r1b}
mkdir -p gpurun_out
for WL in c4 c2 c5; do
  case $WL in c4) K=fused_filter;; c2) K=kmer_scan_kernel;; c5) K=batched_tc_kernel;; esac
  CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-others"
  $CMD > gpurun_out/${TAG}_${WL}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_${WL} $CMD > gpurun_out/${TAG}_${WL}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_${WL}_ncu.log
  $CMD > gpurun_out/${TAG}_${WL}_plain2.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hist_kernel|fused_filter|order_|onehot|profile_|kmer_|batched_|tc_' -c 300 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu1.log 2>&1
done
ls -la gpurun_out | tail -15
