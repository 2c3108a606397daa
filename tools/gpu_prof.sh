#!/bin/bash
# ncu --set full captures of the main kernels of one bench step (config 2). usage: bash tools/gpu_prof.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
export BENCH_SKIP_CPU=1
P="python bench.py --steps 1 --warmup 1"
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:verify_candidates -s 9 -c 2 -o gpurun_out/prof_verify_$TAG $P > gpurun_out/ncu_verify_$TAG.log 2>&1; echo "verify $?"
timeout 600 $N -k regex:reduce_round -s 9 -c 2 -o gpurun_out/prof_reduce_$TAG $P > gpurun_out/ncu_reduce_$TAG.log 2>&1; echo "reduce $?"
timeout 600 $N -k regex:prepare_reads -s 0 -c 1 -o gpurun_out/prof_prepare_$TAG $P > gpurun_out/ncu_prepare_$TAG.log 2>&1; echo "prepare $?"
timeout 600 $N -k regex:seed_lookup -s 9 -c 1 -o gpurun_out/prof_lookup_$TAG $P > gpurun_out/ncu_lookup_$TAG.log 2>&1; echo "lookup $?"
timeout 600 $N -k regex:pair_round -s 0 -c 1 -o gpurun_out/prof_pair_$TAG $P > gpurun_out/ncu_pair_$TAG.log 2>&1; echo "pair $?"
ls -la gpurun_out
