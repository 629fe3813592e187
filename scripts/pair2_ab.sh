#!/bin/bash
# CTA-pair mode (cta_group::2) A/B: per-kernel ms at 2^17 rows and c2
for R in 131072 2500; do for P in 0 1; do
BD_TC_PAIR2=$P timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 8 --warmup 3 --rows $R 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rows=$R pair2=$P', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
"
done; done
