#!/bin/bash
# Run on the GPU box: ncu --set full (with source) of the rpn conv+LIF spike GEMM launch of one step.
# Usage: profiles/ncu_gemm.sh <tag> <mode>
TAG=${1:-r01}
MODE=${2:-fp16x2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 6 -c 1 -o gpurun_out/${TAG}_gemm_${MODE} \
    python bench.py --steps 2 --warmup 3 --mode $MODE --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_gemm_${MODE}.log 2>&1
ls -la gpurun_out | tail -5
