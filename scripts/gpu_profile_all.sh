#!/bin/bash
# Run on the GPU box (under gpurun): every launch of one 60 min ProcessAudio step under `ncu --set full`, reduced to one raw
# CSV (the .ncu-rep of ~270 launches is too big to travel back), plus the launch list of the bench command itself.
# Usage: scripts/gpu_profile_all.sh <tag> [minutes]
TAG=${1:-r2}; MIN=${2:-60}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --minutes $MIN --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none -k regex:k_ -o /tmp/prof_${TAG}_all -f \
    python scripts/profile_step.py --minutes $MIN --flac > gpurun_out/ncu_all_${TAG}.log 2>&1
ncu -i /tmp/prof_${TAG}_all.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_all.csv 2>> gpurun_out/ncu_all_${TAG}.log
ls -la /tmp/prof_${TAG}_all.ncu-rep gpurun_out | tail -8
