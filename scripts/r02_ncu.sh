#!/bin/bash
# usage: scripts/r02_ncu.sh TAG -- ncu --set full captures of the rollout and BPTT kernels at c2 and at 2^17 rows
TAG=$1
mkdir -p gpurun_out
B="python bench.py --no-extra --no-cpu-baseline --steps 1 --warmup 1 --no-graph"
timeout 300 $B > /dev/null 2>&1 || exit 1
for K in bptt_kernel "rollout_fwd_kernel<0, 1, 1"; do
  N=$(echo $K | cut -c1-7)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o gpurun_out/${TAG}_ncu_full_c2_$N $B > gpurun_out/${TAG}_ncu_c2_$N.log 2>&1
  echo c2 $N rc=$?
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o gpurun_out/${TAG}_ncu_full_131k_$N $B --rows 131072 > gpurun_out/${TAG}_ncu_131k_$N.log 2>&1
  echo 131k $N rc=$?
done
