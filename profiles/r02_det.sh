#!/bin/bash
# Run on the GPU box: ncu of the detector post-processing kernel, per-line samples exported as CSV.
TAG=${1:-r02ah}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:det_postprocess -s 5 -c 1 -o gpurun_out/${TAG}_det python profiles/bench_next_rows.py > gpurun_out/${TAG}_det.log 2>&1
F=gpurun_out/${TAG}_det.ncu-rep
ncu -i $F --page raw --csv > gpurun_out/${TAG}_det_raw.csv 2>/dev/null
ncu -i $F --page source --csv --print-source cuda > gpurun_out/${TAG}_det_source_cuda.csv 2>/dev/null
ncu -i $F --page source --csv --print-source sass > gpurun_out/${TAG}_det_source_sass.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_det_source_sass.csv
rm -f $F
ls -la gpurun_out | grep ${TAG}
