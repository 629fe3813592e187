#!/bin/bash
# usage: scripts/quick.sh TAG -- parity tests of the tensor-core path + bench at c2 and 2^17 rows
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_round2.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
for R in 2500 131072; do
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 --rows $R 2>gpurun_out/${TAG}_bench_$R.err | tee gpurun_out/${TAG}_bench_$R.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['start_states_per_gpu'], round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, 'step_frac', round(d['roofline']['step_frac'],4))
"
done
