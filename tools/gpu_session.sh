#!/bin/bash
# One GPU session: parity tests, (on failure: compute-sanitizer over the smoke case), bench line, ncu --set full of chosen kernels.
# usage: bash tools/gpu_session.sh <tag> ["<pytest -k expr>"] ["<kernel-regex>:<skip>:<count> ..."] [bench steps]
TAG=${1:-s}; KEXPR=${2:-}; SPECS=${3:-"screen_bits:9:2 prepare_reads:1:1"}; STEPS=${4:-10}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if [ -n "$KEXPR" ]; then K=(-k "$KEXPR"); else K=(); fi
(time timeout 1500 python -m pytest tests -m gpu -x -q "${K[@]}") > gpurun_out/pytest_$TAG.log 2>&1; PRC=$?
echo "pytest exit $PRC"; tail -25 gpurun_out/pytest_$TAG.log | cut -c1-300
if [ $PRC -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$TAG.log 2>&1; echo "sanitizer exit $?"; grep -E "Invalid|ERROR SUMMARY|at .*align.cu|smoke" gpurun_out/sanitizer_$TAG.log | head -30
fi
BENCH_SKIP_CPU=1 timeout 900 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
export BENCH_SKIP_CPU=1
for spec in $SPECS; do
  IFS=: read -r KN S C <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:$KN -s $S -c $C -o gpurun_out/prof_${KN}_$TAG python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_${KN}_$TAG.log 2>&1; echo "ncu $KN $?"
done
ls -la gpurun_out | tail -6
