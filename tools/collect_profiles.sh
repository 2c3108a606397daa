#!/bin/bash
# Copy the evidence of one tools/gpu_evidence.sh session from gpurun_out/ (scratch) into profiles/ (tracked), round 2 names.
# usage: bash tools/collect_profiles.sh <session tag> [candidates per step] [algorithmic bytes per candidate]
TAG=$1; CAND=${2:-87047216}; BPC=${3:-52}; G=gpurun_out; P=profiles
grep '^{' $G/bench_$TAG.json > $P/bench_r2_final.json
for c in 1 3 4 5; do [ -s $G/bench_${TAG}_c$c.json ] && grep '^{' $G/bench_${TAG}_c$c.json > $P/bench_r2_config$c.json; done
grep -v '^==' $G/launches_$TAG.csv > $P/launches_r2_final.csv
python tools/ncu_summary.py $G/launches_$TAG.csv > $P/launches_r2_final_summary.txt
python tools/ncu_traffic.py $G/prof_verify_step_$TAG.ncu-rep $P/roofline_traffic.json $CAND $BPC > $P/ncu_r2_screen_bits_step.txt
{ python tools/ncu_metrics.py $G/prof_verify_step_$TAG.ncu-rep | head -60; python tools/ncu_lines.py $G/prof_verify_step_$TAG.ncu-rep | head -45; } >> $P/ncu_r2_screen_bits_step.txt 2>&1
{ python tools/ncu_metrics.py $G/prof_prepare_reads_$TAG.ncu-rep; python tools/ncu_lines.py $G/prof_prepare_reads_$TAG.ncu-rep | head -45; } > $P/ncu_r2_prepare_reads.txt 2>&1
ls -la $P | grep r2
