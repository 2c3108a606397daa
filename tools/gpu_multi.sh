#!/bin/bash
# bench.py under torchrun on N GPUs of one box, both arms. usage: bash tools/gpu_multi.sh <tag> <N>
TAG=$1; N=$2; mkdir -p gpurun_out
nvidia-smi -L
BENCH_SKIP_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "exit $?"
tail -3 gpurun_out/bench_${TAG}_n$N.err; cat gpurun_out/bench_${TAG}_n$N.json
