#!/bin/bash
# usage: scripts/r02_final.sh TAG -- full GPU pass for the profiles/ directory: tests, the bench line, the ncu launch
# list of the bench command and full captures of the dominant kernels (each only after the plain run exited 0)
TAG=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_fp16_c2.json 2> gpurun_out/${TAG}_bench.err
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2> gpurun_out/${TAG}_bench_ref.err
echo reference rc=$?
timeout 600 python bench.py --precision fp32 --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_fp32_c2.json 2> gpurun_out/${TAG}_bench_fp32.err
echo fp32 rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_fp16_c2.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 --no-graph > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bptt_kernel|rollout_fwd_kernel|mlp_bwd_kernel' -s 6 -c 6 -o gpurun_out/${TAG}_ncu_full_c2 python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 --no-graph > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
echo full c2 rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bptt_kernel|rollout_fwd_kernel|mlp_bwd_kernel|wgrad_kernel' -s 4 -c 4 -o gpurun_out/${TAG}_ncu_full_131k python bench.py --no-extra --no-cpu-baseline --steps 1 --warmup 1 --no-graph --rows 131072 > gpurun_out/${TAG}_ncu_full_131k.log 2>&1
echo full 131k rc=$?
