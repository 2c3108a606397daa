bash tools/gpu_tests.sh r2x "matches_oracle or mixed" 0
export BENCH_SKIP_CPU=1
for k in 3 4 5 6; do
  BENCH_INFLIGHT=$k timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2x_i$k.json 2> gpurun_out/bench_r2x_i$k.err; echo "inflight $k exit $?"
  python - <<PY
import json
for l in open('gpurun_out/bench_r2x_i$k.json'):
    if l.startswith('{'):
        d=json.loads(l); print('inflight $k value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), [round(x/1e6,1) for x in d['e2e']['repetitions_reads_per_s']], 'pack', round(d['roofline']['ms_pack_per_step'],2))
PY
done
