#!/bin/bash
# A/B of environment settings on the resident replay + x265-host e2e: usage tools/ab.sh <workload> "ENV=.. ENV=.." "ENV=.." ...
W=$1; shift
for cfg in "$@"; do
  env $cfg python bench.py --quick --no-cpu-baseline --no-parity --workload $W 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms_per_step']
print('$W [$cfg]', 'value %.0f (%.2f ms)'%(d['value'], d['ms_per_step']), 'search %.2f cost %.2f intra %.2f'%(k['search'],k['cost'],k['intra']), 'e2e %.0f (%.2f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']))"
done
