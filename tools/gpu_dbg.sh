#!/bin/bash
# A/B of screen_bits variants (BSL_DBG bit mask): a few ncu metrics of the 10th launch. usage: bash tools/gpu_dbg.sh "<dbg values>"
mkdir -p gpurun_out
export BENCH_SKIP_CPU=1
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum
for d in $1; do
  BSL_DBG=$d timeout 600 ncu --metrics $M --clock-control none -k regex:screen_bits -s 9 -c 1 --csv --log-file gpurun_out/dbg_$d.csv python bench.py --steps 1 --warmup 1 > gpurun_out/dbg_$d.log 2>&1
  echo "== dbg $d"; grep -v "^==" gpurun_out/dbg_$d.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done
