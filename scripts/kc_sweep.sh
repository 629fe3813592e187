#!/bin/bash
# usage: scripts/kc_sweep.sh TAG KC...  -- bench at c2 and 2^17 rows for several weight-ring stage depths (BD_TC_KC)
TAG=$1; shift
mkdir -p gpurun_out
for KC in "$@"; do
for R in 2500 131072; do
BD_TC_KC=$KC timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 --rows $R 2>gpurun_out/${TAG}_kc${KC}_$R.err | tee gpurun_out/${TAG}_kc${KC}_$R.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('KC=$KC', d['config']['start_states_per_gpu'], round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, 'step_frac', round(d['roofline']['step_frac'],4))
"
done
done
