#!/bin/bash
# parity tests only (+ optional full bench line with cpu_baseline / parity_check / cli_e2e). usage: bash tools/gpu_tests.sh <tag> ["<pytest -k>"] [bench: 0|1]
TAG=${1:-t}; KEXPR=${2:-}; BENCH=${3:-0}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then K=(-k "$KEXPR"); else K=(); fi
(time timeout 2400 python -m pytest tests -m gpu -q "${K[@]}") > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_$TAG.log | cut -c1-400
if [ "$BENCH" = "1" ]; then
  timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
fi
