#!/bin/bash
# ncu captures of the kernels that changed late in round 2: batched_tc_kernel with the sequence mask, filter_q4_kernel
mkdir -p gpurun_out
CMD="python bench.py --workload c5 --n-per-gpu 125000000 --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --sustain-seconds 0"
$CMD > gpurun_out/r2n_c5_plain.json 2> gpurun_out/r2n_c5_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:batched_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2n_c5 $CMD > gpurun_out/r2n_c5_ncu.log 2>&1
tail -1 gpurun_out/r2n_c5_ncu.log
E2E="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-sweep --sustain-seconds 0 --no-api --n-per-gpu 125000000 --e2e-form q4"
python bench.py $E2E > gpurun_out/r2n_e2e_q4_plain.json 2> gpurun_out/r2n_e2e_q4_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_q|resolve_kernel|refine_packed|order_kernel' -c 400 --csv \
    --log-file gpurun_out/r2n_e2e_q4_launches.csv python bench.py $E2E > gpurun_out/r2n_e2e_q4_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_q4 -s 20 -c 1 -f -o gpurun_out/r2n_q4 \
    python bench.py $E2E > gpurun_out/r2n_q4_ncu.log 2>&1
tail -1 gpurun_out/r2n_q4_ncu.log
