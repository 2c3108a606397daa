#!/bin/bash
# quick GPU check: parity tests + bench line (no CPU baseline) + launch list. usage: bash tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${2:-} > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
