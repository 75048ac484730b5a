#!/bin/bash
# Wall time of the CLI itself, one process vs one rank per GPU (run with gpurun --gpus N).
N=${1:-8}
D=/dev/shm/cli_scale
mkdir -p gpurun_out
python tools/cli_scale.py $D ${2:-200e6} ${3:-50e6} 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
OUT=gpurun_out/cli_scale_n${N}.jsonl
rm -f $OUT gpurun_out/cli_scale_n${N}_wall.txt
run() {   # name, launcher..., then the rnascan arguments
  name=$1; shift
  t0=$(date +%s.%N)
  "$@" --stats $OUT > /dev/shm/hits_$name.tab 2> /dev/shm/err_$name.log
  rc=$?
  t1=$(date +%s.%N)
  echo "$name rc=$rc wall $(python -c "print('%.2f' % ($t1 - $t0))") s" | tee -a gpurun_out/cli_scale_n${N}_wall.txt
  grep -v "^\*\|OMP_NUM" /dev/shm/err_$name.log | tail -2
}
RNA="-p $D/seq.pfm $D/seqs.fa"
SS="-q $D/struct.pfm -B $D/bg_struct.txt $D/profiles"
for W in 1 $N; do
  if [ $W -eq 1 ]; then L="python -m rnascan_b200.rnascan"; else L="$TR -m rnascan_b200.rnascan"; fi
  run rna_w$W $L $RNA
  run ss_w$W $L $SS
done
cmp /dev/shm/hits_rna_w1.tab /dev/shm/hits_rna_w$N.tab && cmp /dev/shm/hits_ss_w1.tab /dev/shm/hits_ss_w$N.tab && echo "outputs identical at 1 and $N ranks"
wc -l /dev/shm/hits_rna_w1.tab /dev/shm/hits_ss_w1.tab
cat $OUT
rm -rf $D
