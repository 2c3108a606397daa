#!/bin/bash
# one step's verify_candidates launches under ncu --set full (for roofline.traffic) + launch list. usage: bash tools/gpu_traffic.sh <tag> <n_launches>
TAG=$1; N=${2:-36}; mkdir -p gpurun_out
export BENCH_SKIP_CPU=1
timeout 1200 ncu --set full --clock-control none --import-source on -f -k regex:verify_candidates -s 0 -c $N -o gpurun_out/prof_verify_step_$TAG python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_verify_step_$TAG.log 2>&1; echo "ncu $?"
bash tools/gpu_launches.sh $TAG
