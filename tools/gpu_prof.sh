#!/bin/bash
# ncu --set full captures of chosen kernels of one bench step (config 2).
# usage: bash tools/gpu_prof.sh <tag> "<kernel-regex>:<skip>:<count> ..."   [bench first: set BENCH_FIRST=1]
TAG=${1:-x}; shift
SPECS=${1:-"screen_bits:0:2 reduce_round:0:2 prepare_reads:1:1"}
mkdir -p gpurun_out
export BENCH_SKIP_CPU=1
if [ -n "$BENCH_FIRST" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json
fi
P="python bench.py --steps 1 --warmup 1"
for spec in $SPECS; do
  IFS=: read -r K S C <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:$K -s $S -c $C -o gpurun_out/prof_${K}_$TAG $P > gpurun_out/ncu_${K}_$TAG.log 2>&1; echo "$K $?"
done
ls -la gpurun_out | tail -8
