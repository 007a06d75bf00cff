#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the heaviest kernels.
# Usage: scripts/gpu_profile.sh <tag> [minutes] [kernel regexes...]
TAG=${1:-r1}; MIN=${2:-10}; shift 2
KERNELS=${@:-k_dc_interp k_anlmdn k_afftdn_fwd}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --minutes $MIN --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/prof_${TAG}_$K -f \
      python bench.py --minutes $MIN --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
ls -la gpurun_out | tail -12
