#!/bin/bash
# usage: scripts/gpu_check.sh TAG  -- quick parity + timing pass of the tensor-core path (writes gpurun_out/TAG_*)
TAG=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_tc.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
N=2560 BD_TC_CLUSTER=1 timeout 120 python scripts/prof_fwd.py > gpurun_out/${TAG}_time_fwd_c2_r1.log 2>&1
N=2560 timeout 120 python scripts/prof_fwd.py > gpurun_out/${TAG}_time_fwd_c2.log 2>&1
N=18944 timeout 120 python scripts/prof_fwd.py > gpurun_out/${TAG}_time_fwd_148.log 2>&1
N=18944 BD_TC_PROF=1 timeout 120 python scripts/prof_fwd.py > gpurun_out/${TAG}_prof_fwd_148.log 2>&1
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 --rows 131072 > gpurun_out/${TAG}_bench_131k.json 2> gpurun_out/${TAG}_bench_131k.err
tail -3 gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_time_fwd_*.log
