#!/bin/bash
TAG=${1:-x}; mkdir -p gpurun_out
BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
