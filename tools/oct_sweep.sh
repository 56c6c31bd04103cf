#!/bin/bash
# quick A/B of the octet kernel's wait parameters (resident replay only): prints search ms per step
for cfg in "0 64 64" "4 32 256" "8 32 256" "8 32 1000" "16 32 500"; do
  set -- $cfg
  X265CU_OCT_SLACK=$1 X265CU_OCT_SLEEP=$2 X265CU_OCT_SLEEP_FULL=$3 python bench.py --quick --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('slack/sleep/sleepfull $cfg', 'value %.0f'%d['value'], 'search %.2f ms'%d['kernel_ms_per_step']['search'], 'e2e %.0f'%d['e2e']['value'])"
done
X265CU_PLAIN_OCT=0 python bench.py --quick --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('quad kernel', 'value %.0f'%d['value'], 'search %.2f ms'%d['kernel_ms_per_step']['search'], 'e2e %.0f'%d['e2e']['value'])"
