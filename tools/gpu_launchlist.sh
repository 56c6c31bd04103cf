#!/bin/bash
# launch list (device time of every launch) of bench.py --profile-mode; usage: gpu_launchlist.sh <workload> <tag>
W=$1; T=$2; O=gpurun_out; mkdir -p $O
CMD="python bench.py --profile-mode --workload $W"
$CMD > $O/${T}_plain.json 2> $O/${T}_plain.err || { echo "plain run failed"; tail -5 $O/${T}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_$W.csv $CMD > $O/${T}_ncu_list.log 2>&1
tail -2 $O/${T}_ncu_list.log
