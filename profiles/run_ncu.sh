#!/bin/bash
# Run on the GPU box (via gpurun).  Usage: profiles/run_ncu.sh <tag> [mode]
# Produces gpurun_out/<tag>_launches.csv (every launch, device time) and
# gpurun_out/<tag>_gemm.ncu-rep (--set full of the spike GEMM launches of one step).
TAG=${1:-r01}
MODE=${2:-fp32_exact}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-e2e --no-cpu-baseline"
# bench warms up 3 steps (17 launches each + 2 weight-prep) before the 2 timed steps
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 9 -c 3 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_|readout_" -s 24 -c 12 -o gpurun_out/${TAG}_aux $CMD > gpurun_out/${TAG}_aux.log 2>&1
ls -la gpurun_out
