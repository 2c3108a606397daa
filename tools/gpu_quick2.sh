#!/bin/bash
# parity subset + bench + optional profiles. usage: bash tools/gpu_quick2.sh <tag> "<pytest -k expr>" "<prof specs>"
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "${2:-matches_oracle}" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
BENCH_FIRST=1 bash tools/gpu_prof.sh $TAG "${3:-screen_bits:9:1}"
