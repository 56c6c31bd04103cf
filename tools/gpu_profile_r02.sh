#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): launch lists of `bench.py --profile-mode` (device-resident replay of one
# step) and `--set full` captures of one launch of every kernel family at 4K.  Summaries: tools/ncu_summary.py on the CPU box.
# usage: tools/gpu_profile_r02.sh <tag>
T=${1:-r02k}
O=gpurun_out
mkdir -p $O
for W in c2_4k c1_1080p; do
  CMD="python bench.py --profile-mode --workload $W"
  $CMD > $O/${T}_plain_$W.json 2> $O/${T}_plain_$W.err || { echo "plain run failed"; tail -5 $O/${T}_plain_$W.err; exit 1; }
  cat $O/${T}_plain_$W.json
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_$W.csv $CMD > $O/${T}_ncu_list_$W.log 2>&1
done
W=c2_4k
CMD="python bench.py --profile-mode --workload $W"
for K in lowres_init_batch_kernel frame_var_batch_kernel intra_batch_kernel cutree_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o $O/${T}_${W}_$K $CMD > $O/${T}_ncu_$K.log 2>&1
  tail -1 $O/${T}_ncu_$K.log
done
ncu --set full --clock-control none --import-source on -k regex:plain_search_kernel -c 2 -f -o $O/${T}_${W}_plain_search_kernel $CMD > $O/${T}_ncu_search.log 2>&1
tail -1 $O/${T}_ncu_search.log
ncu --set full --clock-control none --import-source on -k regex:'^cost_kernel|void cost_kernel' -c 2 -f -o $O/${T}_${W}_cost_kernel $CMD > $O/${T}_ncu_cost.log 2>&1
tail -1 $O/${T}_ncu_cost.log
# whole-frame SATD primitive: the first cold launch of the wide kernel (after 6 warm ones) of tools/satd_bw.py
ncu --set full --clock-control none --import-source on -k regex:pixelcmp_frames_wide_kernel --launch-skip 6 -c 1 -f -o $O/${T}_${W}_pixelcmp_wide python tools/satd_bw.py $W > $O/${T}_ncu_pixel.log 2>&1
tail -1 $O/${T}_ncu_pixel.log
python tools/pcie_bw.py > $O/${T}_pcie.txt 2>&1; tail -8 $O/${T}_pcie.txt
# summaries are made here (gpurun brings back at most 64 MiB): every launch of every report, then only the smaller reports stay
for R in $O/${T}_${W}_*.ncu-rep; do
  N=$(ncu -i $R --page raw --csv 2>/dev/null | tail -n +3 | wc -l)
  for ((i = 0; i < N; i++)); do
    echo "=== $(basename $R .ncu-rep) launch $i" >> $O/${T}_ncu_summary.txt
    python tools/ncu_summary.py $R $i >> $O/${T}_ncu_summary.txt 2>&1
  done
done
rm -f $O/${T}_${W}_cost_kernel.ncu-rep $O/${T}_${W}_cutree_kernel.ncu-rep $O/${T}_${W}_intra_batch_kernel.ncu-rep $O/${T}_${W}_pixelcmp_wide.ncu-rep
ls -la $O | grep ${T}_
