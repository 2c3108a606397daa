export BENCH_SKIP_CPU=1
run() { env $1 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r3e_$2.json 2> gpurun_out/bench_r3e_$2.err
python - <<PY
import json
for l in open('gpurun_out/bench_r3e_$2.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$2', 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k:round(v,3) for k,v in r.items() if k.startswith('ms_')})
PY
}
run X=0 base
run BSL_X_WIDE=45 wide45
run BSL_X_RR=50 rr50
run BSL_X_SB=50 sb50
run BSL_X_SB=100 sb100
