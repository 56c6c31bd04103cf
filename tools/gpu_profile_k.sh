#!/bin/bash
# ncu --set full of the first launches of one kernel of `bench.py --profile-mode` (run under gpurun)
# usage: tools/gpu_profile_k.sh <workload> <tag> <kernel regex> [count] [extra env assignments...]
W=$1; T=$2; K=$3; N=${4:-2}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --profile-mode --workload $W"
$CMD > $O/${T}_plain.json 2> $O/${T}_plain.err || { echo "plain run failed"; tail -5 $O/${T}_plain.err; exit 1; }
cat $O/${T}_plain.json
ncu --set full --clock-control none --import-source on -k regex:$K -c $N -f -o $O/${T}_${W} $CMD > $O/${T}_ncu.log 2>&1
tail -3 $O/${T}_ncu.log
