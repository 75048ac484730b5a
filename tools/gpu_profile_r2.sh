#!/bin/bash
# Round-2 evidence in one gpurun call: plain runs first, then ncu (launch lists + one full capture per headline kernel).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
FAST="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --no-sweep --sustain-seconds 0"
# ---- C4 at 1 Gnt (the default workload): full capture of the fused filter kernel + launch list
python bench.py $FAST > gpurun_out/${TAG}_c4_plain.json 2> gpurun_out/${TAG}_c4_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:fused_filter -s 3 -c 1 -f -o gpurun_out/${TAG}_c4 \
    python bench.py $FAST > gpurun_out/${TAG}_c4_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_c4_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
    python bench.py $FAST > gpurun_out/${TAG}_c4_ncu1.log 2>&1
# ---- C4 end to end (quantised rows): launch list of the HostProfileScanner pipeline at 125 M symbols
E2E="--steps 2 --warmup 3 --no-cpu-baseline --no-others --no-sweep --sustain-seconds 0 --no-api --n-per-gpu 125000000"
python bench.py $E2E > gpurun_out/${TAG}_e2e_plain.json 2> gpurun_out/${TAG}_e2e_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_q8|resolve_kernel|order_kernel|hist' -c 300 --csv \
    --log-file gpurun_out/${TAG}_e2e_launches.csv python bench.py $E2E > gpurun_out/${TAG}_e2e_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_q8 -s 20 -c 1 -f -o gpurun_out/${TAG}_q8 \
    python bench.py $E2E > gpurun_out/${TAG}_q8_ncu.log 2>&1
# ---- C5 / C2 / C3 at 125 M: full capture of the dominant kernel + launch list
for WL in c5 c2 c3; do
  case $WL in c5) K=batched_tc_kernel;; c2) K=kmer_scan_kernel;; c3) K=dense_w_kernel;; esac
  CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-others --no-e2e --sustain-seconds 0"
  $CMD > gpurun_out/${TAG}_${WL}_plain.json 2> gpurun_out/${TAG}_${WL}_plain.err &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_${WL} $CMD > gpurun_out/${TAG}_${WL}_ncu.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu1.log 2>&1
done
# ---- the CLI itself: text profile directory + FASTA, combined mode, first with text then with the pack
python tools/cli_dataset.py /tmp/ds 3e6 1000 > /tmp/argv.txt
ARGS=$(tr "\n" " " < /tmp/argv.txt)
python -m rnascan_b200.rnascan $ARGS -m 2 --pack --stats gpurun_out/${TAG}_cli_stats.json > /tmp/hits_text.tab 2> gpurun_out/${TAG}_cli_plain.log
python -m rnascan_b200.rnascan $ARGS -m 2 --stats gpurun_out/${TAG}_cli_stats.json > /tmp/hits_pack.tab 2>> gpurun_out/${TAG}_cli_plain.log
cmp /tmp/hits_text.tab /tmp/hits_pack.tab && wc -l /tmp/hits_text.tab >> gpurun_out/${TAG}_cli_plain.log
rm -f /tmp/ds/profiles/rnascan_b200.pack
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cli_text_launches.csv \
    python -m rnascan_b200.rnascan $ARGS -m 2 --pack > /tmp/h1.tab 2> gpurun_out/${TAG}_cli_ncu_text.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cli_pack_launches.csv \
    python -m rnascan_b200.rnascan $ARGS -m 2 > /tmp/h2.tab 2> gpurun_out/${TAG}_cli_ncu_pack.log
cmp /tmp/h1.tab /tmp/hits_text.tab && cmp /tmp/h2.tab /tmp/hits_text.tab && echo "CLI outputs identical (text, pack, under ncu)" >> gpurun_out/${TAG}_cli_plain.log
ls -la gpurun_out | tail -40
cat gpurun_out/${TAG}_cli_plain.log | tail -12
cat gpurun_out/${TAG}_cli_stats.json
