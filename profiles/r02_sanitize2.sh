#!/bin/bash
# Run on the GPU box: compute-sanitizer over the kernels that changed after r02f (radix-select top-k, the lean fc producers
# with word reuse, the conv producers' sleeping wait, the weight-multicast schedule).  Usage: profiles/r02_sanitize2.sh <tag>
TAG=${1:-r02z}
mkdir -p gpurun_out
SEL="topk_ties or rpn_proposals_from_native or box_fp32_exact_vs_reference_golden or box_without_recording or multicast_schedule or fused_roi_align"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "topk_ties or box_without_recording" > gpurun_out/${TAG}_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/${TAG}_synccheck.log; tail -4 gpurun_out/${TAG}_synccheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests -m gpu -q -x -k "topk_ties" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log; grep -E "RACECHECK SUMMARY|hazard|Race reported" gpurun_out/${TAG}_racecheck.log | cut -c1-200 | sort | uniq -c | head -12; tail -3 gpurun_out/${TAG}_racecheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "topk_ties or box_without_recording" > gpurun_out/${TAG}_initcheck.log 2>&1
echo "initcheck rc=$?" >> gpurun_out/${TAG}_initcheck.log; tail -4 gpurun_out/${TAG}_initcheck.log
