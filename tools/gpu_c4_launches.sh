#!/bin/bash
# launch list (ncu, our kernels only) of the default workload's step as it is at the end of the round
mkdir -p gpurun_out
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --no-sweep --no-api --sustain-seconds 0"
timeout 300 python bench.py $FAST > gpurun_out/r2s_c4_plain.json 2> gpurun_out/r2s_c4_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fused_filter|order_kernel|hist_kernel|refine|resolve' -c 60 --csv \
    --log-file gpurun_out/r2s_c4_launches.csv python bench.py $FAST > gpurun_out/r2s_c4_ncu1.log 2>&1
tail -6 gpurun_out/r2s_c4_launches.csv | cut -c1-260
