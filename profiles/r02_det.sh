#!/bin/bash
# Run on the GPU box: ncu of the two post-processing kernels (SURVEY 8f-1 tail, 8f-4), one launch each, raw metrics and
# per-instruction samples exported as CSV.
TAG=${1:-r02as}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"det_postprocess|rpn_nms" -s 6 -c 2 -o gpurun_out/${TAG}_post python profiles/bench_next_rows.py > gpurun_out/${TAG}_post.log 2>&1
F=gpurun_out/${TAG}_post.ncu-rep
ncu -i $F --page raw --csv > gpurun_out/${TAG}_post_raw.csv 2>/dev/null
ncu -i $F --page source --csv --print-source sass > gpurun_out/${TAG}_post_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_post_source.csv
rm -f $F
ls -la gpurun_out | grep ${TAG}
