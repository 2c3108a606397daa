bash tools/gpu_tests.sh r3l "matches_oracle or mixed" 0
export BENCH_SKIP_CPU=1
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:screen_bits -s 0 -c 9 -o gpurun_out/prof_verify_step_r3l python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_verify_step_r3l.log 2>&1; echo "ncu step $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r3l.json 2> gpurun_out/bench_r3l.err; echo "bench $?"; cut -c1-200 gpurun_out/bench_r3l.json
