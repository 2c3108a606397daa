#!/bin/bash
# A/B bench runs with different environment settings. usage: bash tools/gpu_ab.sh <tag> "ENV1=.." "ENV2=.." ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  i=$((i+1))
  env $e BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  echo "== $e"; python - <<PY
import json; d=json.load(open('gpurun_out/bench_${TAG}_$i.json')); print(round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline'].items() if k.startswith('ms_') or k in ('frac',)})
PY
done
