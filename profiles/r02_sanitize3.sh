#!/bin/bash
# Run on the GPU box: compute-sanitizer over the detector post-processing and proposal-filter kernels (snn_det_postprocess, snn_rpn_nms).
TAG=${1:-r02aj}
mkdir -p gpurun_out
SEL="postprocess_kernel or rpn_filter_kernel"
for TOOL in memcheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 1 python -m pytest tests/test_detection_post.py -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_${TOOL}.log 2>&1
  echo "$TOOL rc=$?" >> gpurun_out/${TAG}_${TOOL}.log; tail -4 gpurun_out/${TAG}_${TOOL}.log
done
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_detection_post.py -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log; grep -E "RACECHECK SUMMARY|Race reported|hazard" gpurun_out/${TAG}_racecheck.log | cut -c1-220 | sort | uniq -c | head -12; tail -3 gpurun_out/${TAG}_racecheck.log
