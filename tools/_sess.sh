export BENCH_SKIP_CPU=1
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r3f.json 2> gpurun_out/bench_r3f.err
python - <<PY
import json
for l in open('gpurun_out/bench_r3f.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ldg', 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k:round(v,3) for k,v in r.items() if k.startswith('ms_')})
PY
