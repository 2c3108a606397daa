#!/bin/bash
# Round evidence in one GPU session: bench line (cpu_baseline, parity_check, cli_e2e), ncu launch list of one step, ncu --set full
# of one step's screen_bits launches (roofline.traffic) and of prepare_reads, bench lines of the other configs.
# usage: bash tools/gpu_evidence.sh <tag> [configs]
TAG=${1:-ev}; CFGS=${2:-"1 3 4 5"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_$TAG.json
export BENCH_SKIP_CPU=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:screen_bits -s 0 -c 9 -o gpurun_out/prof_verify_step_$TAG python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_verify_step_$TAG.log 2>&1; echo "ncu step $?"
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:prepare_reads -s 1 -c 1 -o gpurun_out/prof_prepare_reads_$TAG python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_prepare_$TAG.log 2>&1; echo "ncu prepare $?"
for c in $CFGS; do
  BENCH_CONFIG=$c timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "config $c exit $?"; cut -c1-300 gpurun_out/bench_${TAG}_c$c.json
done
ls -la gpurun_out | tail -12
