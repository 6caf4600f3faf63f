#!/bin/bash
# Run on the GPU box: ncu --set full (with source) of ONE spike GEMM launch.  Usage: profiles/ncu_one.sh <tag> <mode> <skip>
# skip = index of the launch among the spike_gemm_lif launches (3 per step: conv, fc6, fc7; 3 warm-up steps first)
TAG=${1:-r01}; MODE=${2:-fp16x2}; SKIP=${3:-10}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s $SKIP -c 1 -o gpurun_out/${TAG}_one_${MODE}_${SKIP} \
    python bench.py --steps 2 --warmup 3 --mode $MODE --no-e2e --no-cpu-baseline --no-other-modes > gpurun_out/${TAG}_one_${MODE}_${SKIP}.log 2>&1
