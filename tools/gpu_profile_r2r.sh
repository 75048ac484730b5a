#!/bin/bash
# Final captures of the round: the fused filter kernel as it now runs (counting the background itself), the C2 chain
# (launch list: no memset nodes, histogram through dp4a) and the histogram kernel.
set -u
TAG=r2r
mkdir -p gpurun_out
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --no-sweep --no-api --sustain-seconds 0"
timeout 300 python bench.py $FAST > gpurun_out/${TAG}_c4_plain.json 2> gpurun_out/${TAG}_c4_plain.err &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_filter -s 3 -c 1 -f -o gpurun_out/${TAG}_c4 \
    python bench.py $FAST > gpurun_out/${TAG}_c4_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_c4_ncu.log
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --sustain-seconds 0"
timeout 300 $CMD > gpurun_out/${TAG}_c2_plain.json 2> gpurun_out/${TAG}_c2_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hist_kernel|kmer_|counts_notify' -c 100 --csv \
    --log-file gpurun_out/${TAG}_c2_launches.csv $CMD > gpurun_out/${TAG}_c2_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_hist \
    $CMD > gpurun_out/${TAG}_hist_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_c2_launches.csv
ls -la gpurun_out/${TAG}_*.ncu-rep
