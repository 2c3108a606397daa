#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the roofline kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
  tail -5 gpurun_out/pytest_$TAG.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
BENCH_SKIP_CPU=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:screen_bits -s 9 -c 3 -o gpurun_out/prof_verify_$TAG -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
