#!/bin/bash
# e2e time of the x265-hosted lookahead for several settings of the speculative path's admission rule
for w in ${WORKLOADS:-c1_1080p c2_4k}; do
for plans in 4 16 64; do for dist in 2 4 8; do
  echo -n "$w plans=$plans dist=$dist: "
  X265CU_SPEC_MAX_PLANS=$plans X265CU_SPEC_MAX_DIST=$dist X265CU_GLUE_PROFILE=1 python tools/gpuhost_time.py $w 3 gpu 2>&1 | grep -E "run 2|x265glue" | tail -2 | sed -e 's/x265glue: //' | tr '\n' ' '
  echo
done; done; done
