#!/bin/bash
# Wall time of single-process CLI runs (start-up included), each run twice; --stats shows the phases.
D=/dev/shm/cli_wall
mkdir -p gpurun_out
python tools/cli_scale.py $D ${1:-50e6} ${2:-10e6} 2>&1 | tail -1
OUT=gpurun_out/cli_wall.jsonl
rm -f $OUT gpurun_out/cli_wall.txt
run() {
  name=$1; shift
  t0=$(date +%s.%N)
  python -m rnascan_b200.rnascan "$@" --stats $OUT > /dev/shm/hits_$name.tab 2> /dev/shm/err_$name.log
  rc=$?
  t1=$(date +%s.%N)
  echo "$name rc=$rc wall $(python -c "print('%.2f' % ($t1 - $t0))") s, $(wc -l < /dev/shm/hits_$name.tab) lines" | tee -a gpurun_out/cli_wall.txt
}
for i in 1 2; do
  run rna_$i -p $D/seq.pfm $D/seqs.fa
  run ss_$i -q $D/struct.pfm -B $D/bg_struct.txt $D/profiles
  run rnass_$i -p $D/seq.pfm -q $D/struct.pfm -B $D/bg_struct.txt $D/seqs.fa $D/profiles -m 2
done
python - <<'PY'
import json
for l in open("gpurun_out/cli_wall.jsonl"):
    d = json.loads(l)
    print(d["mode"], round(d["total_s"], 2), {k: round(v, 3) for k, v in d["phases_s"].items()})
PY
rm -rf $D
