bash tools/gpu_tests.sh r2v "matches_oracle or mixed" 0
export BENCH_SKIP_CPU=1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_r2v.json
export BENCH_CONFIG=4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2v_c4.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch_r2v_c4.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -f -k "regex:screen_bits|reduce_round" -s 12 -c 4 -o gpurun_out/prof_c4_r2v python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_c4_r2v.log 2>&1; echo "ncu c4 $?"
