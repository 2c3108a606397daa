#!/bin/bash
# profile kernels of another config: bash tools/gpu_prof_cfg.sh <tag> <config> "<specs>"
TAG=$1; export BENCH_CONFIG=$2; shift; shift
bash tools/gpu_prof.sh $TAG "$1"
