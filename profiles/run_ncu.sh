#!/bin/bash
# Run on the GPU box (via gpurun).  Usage: profiles/run_ncu.sh <tag> [mode]
# Produces in gpurun_out/:
#   <tag>_launches_<mode>.csv   every kernel launch of `bench.py --steps 2 --warmup 3` with its device time
#   <tag>_gemm_<mode>.ncu-rep   --set full (+source) of the four spike GEMM launches of one step (conv, fc6 dual tiles, fc6 tail wave, fc7)
#   <tag>_aux_<mode>.ncu-rep    --set full of the encoder / readout / LUT launches of one step
# Summaries for profiles/ are produced here afterwards with profiles/summarise_ncu.py.
TAG=${1:-r01}
MODE=${2:-fp16x2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --mode $MODE --no-e2e --no-cpu-baseline --no-other-modes"
# per step: rpn {encoder, conv gemm} + box {lut, encoder, fc6 gemm (dual tiles), fc6 gemm (tail wave), fc7 gemm, readout} = 8 launches
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_${MODE}.csv $CMD > gpurun_out/${TAG}_launches_${MODE}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 12 -c 4 -o gpurun_out/${TAG}_gemm_${MODE} $CMD > gpurun_out/${TAG}_gemm_${MODE}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_|readout_|build_lut" -s 12 -c 4 -o gpurun_out/${TAG}_aux_${MODE} $CMD > gpurun_out/${TAG}_aux_${MODE}.log 2>&1
ls -la gpurun_out | tail -8
