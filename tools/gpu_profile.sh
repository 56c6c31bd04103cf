#!/bin/bash
# ncu evidence of one workload (run under gpurun): launch list of `bench.py --profile-mode`, then `--set full` captures
# of the first launches of the dominant kernels.  usage: tools/gpu_profile.sh <workload> <tag>
# Outputs land in gpurun_out/<tag>_*; summaries are made on the CPU box with tools/ncu_summary.py.
W=${1:-c2_4k}
T=${2:-r02}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --profile-mode --workload $W"
$CMD > $O/${T}_plain.json 2> $O/${T}_plain.err || { echo "plain run failed"; tail -5 $O/${T}_plain.err; exit 1; }
cat $O/${T}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_$W.csv $CMD > $O/${T}_ncu_list.log 2>&1
for K in plain_search_kernel cost_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 3 -f -o $O/${T}_${W}_$K $CMD > $O/${T}_ncu_$K.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'lowres_init_kernel|frame_var_kernel|intra_kernel' -c 3 -f -o $O/${T}_${W}_pre $CMD > $O/${T}_ncu_pre.log 2>&1
ls -la $O | tail -20
