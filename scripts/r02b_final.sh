#!/bin/bash
# usage: scripts/r02b_final.sh TAG -- full GPU pass for profiles/: tests, the bench line, --impl reference, fp32 check
# mode, the ncu launch list of the bench command, full captures of the kernels changed in this session
TAG=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
tail -2 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_fp16_c2.json 2> gpurun_out/${TAG}_bench.err
echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2> gpurun_out/${TAG}_bench_ref.err
echo reference rc=$?
timeout 600 python bench.py --precision fp32 --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_fp32_c2.json 2> gpurun_out/${TAG}_bench_fp32.err
echo fp32 rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_fp16_c2.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 --no-graph > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo launches rc=$?
B="python bench.py --no-extra --no-cpu-baseline --steps 1 --warmup 1 --no-graph"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bptt_kernel|rollout_fwd_kernel|mlp_bwd_kernel|wgrad_kernel|actor_entropy' -s 8 -c 8 -o gpurun_out/${TAG}_ncu_full_c2 $B > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
echo full c2 rc=$?
ncu -i gpurun_out/${TAG}_ncu_full_c2.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_c2_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_ncu_full_c2.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bptt_kernel|rollout_fwd_kernel|mlp_bwd_kernel|wgrad_kernel|actor_entropy' -s 5 -c 5 -o gpurun_out/${TAG}_ncu_full_131k $B --rows 131072 > gpurun_out/${TAG}_ncu_full_131k.log 2>&1
echo full 131k rc=$?
ncu -i gpurun_out/${TAG}_ncu_full_131k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_131k_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_ncu_full_131k.ncu-rep
# (the .ncu-rep files together exceed gpurun's 64 MiB merge limit: only the raw pages travel back)
ls -la gpurun_out/${TAG}_*
