bash tools/gpu_tests.sh r2t "" 0
export BENCH_SKIP_CPU=1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_r2t.json
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:prepare_reads -s 1 -c 1 -o gpurun_out/prof_prepare_reads_r2t python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_prepare_r2t.log 2>&1; echo "ncu prepare $?"
