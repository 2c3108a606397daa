#!/bin/bash
# N ranks under torchrun (one per GPU): bench line with scaling, e2e, replica check (and the config-5 block at N = 8); plus the
# 2-GPU test of the CLI. usage: gpurun --gpus N -- bash tools/gpu_multi.sh <tag> <N> [steps]
TAG=${1:-m}; N=${2:-2}; STEPS=${3:-10}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1; nproc >> gpurun_out/topo_${TAG}.txt; free -g | head -2 >> gpurun_out/topo_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_${TAG}.log
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench exit $?"
cat gpurun_out/bench_${TAG}_n$N.json | cut -c1-3000; grep -v "^\[bench\]" gpurun_out/bench_${TAG}_n$N.err | tail -5
