export BENCH_SKIP_CPU=1
for c in 4 1 3; do
BENCH_CONFIG=$c timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_${TAGX}_c$c.json 2> gpurun_out/bench_${TAGX}_c$c.err; echo "config $c exit $?"; cut -c1-200 gpurun_out/bench_${TAGX}_c$c.json
done
