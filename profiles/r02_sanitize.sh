#!/bin/bash
# Run on the GPU box: compute-sanitizer over the kernels touched in round 2 (readout warp + named barriers in the conv,
# box head below three steps, top-k keys, LI statistics, RoIAlign unroll variants).  Usage: profiles/r02_sanitize.sh <tag>
TAG=${1:-r02f}
mkdir -p gpurun_out
SEL="rpn_fp32_exact_vs_reference_golden or below_three or descriptor_cache or spike_rate_report or topk_ties or reduced_weight"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "rpn_fp32_exact_vs_reference_golden or reduced_weight" > gpurun_out/${TAG}_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/${TAG}_synccheck.log; tail -4 gpurun_out/${TAG}_synccheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests -m gpu -q -x -k "rpn_reduced_weight_modes" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_racecheck.log | sort | uniq -c | head -12; tail -3 gpurun_out/${TAG}_racecheck.log
