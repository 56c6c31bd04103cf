#!/bin/bash
for oct in 0 1; do
X265CU_PLAIN_OCT=$oct X265CU_OCT_SLACK=0 X265CU_OCT_SLEEP_FULL=64 python bench.py --configs "" --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('oct=$oct', 'value %.0f'%d['value'], 'search %.2f ms'%d['kernel_ms_per_step']['search'], 'e2e %.0f'%d['e2e']['value'], 'multi8 %.0f'%d['multi_stream']['value'])"
done
