#!/bin/bash
# Round evidence in one GPU session: parity tests, bench line (with cpu_baseline), ncu launch list, ncu --set full of one
# step's screen_bits launches (roofline.traffic), bench lines of the other configs. usage: bash tools/gpu_evidence.sh <tag> [configs]
TAG=${1:-ev}; CFGS=${2:-"1 3 4 5"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_$TAG.log 2>&1; tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json
bash tools/gpu_launches.sh $TAG
bash tools/gpu_traffic.sh $TAG 36 screen_bits
for c in $CFGS; do
  BENCH_CONFIG=$c BENCH_SKIP_CPU=1 timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "config $c exit $?"; cut -c1-400 gpurun_out/bench_${TAG}_c$c.json
done
ls -la gpurun_out | tail -12
