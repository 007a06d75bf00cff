#!/bin/bash
# usage: ncu_one.sh <tag> <kernel regex> <spec> [minutes] [rate] [dtype]   -- one full-set capture with source of one kernel of one filter spec
TAG=$1; K=$2; SPEC=$3; MIN=${4:-10}; RATE=${5:-48000}; DT=${6:-f32}
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/prof_${TAG} -f \
    python scripts/time_filter.py "$SPEC" $MIN $RATE $DT > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
