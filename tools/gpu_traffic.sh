#!/bin/bash
# one step's candidate-verification launches under ncu --set full (for roofline.traffic). usage: bash tools/gpu_traffic.sh <tag> <n_launches = 9 for a paired step> [kernel regex]
TAG=$1; N=${2:-9}; K=${3:-screen_bits}; mkdir -p gpurun_out
export BENCH_SKIP_CPU=1
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:$K -s 0 -c $N -o gpurun_out/prof_verify_step_$TAG python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_verify_step_$TAG.log 2>&1; echo "ncu $?"
