#!/bin/bash
# A/B on one box: fc producers with / without word reuse across a pair's steps (SNN_DBG_PROD_REUSE).  Run r02x also swept a
# sleep between the fc producers' polls (0 / 100 / 250 ns: costs the bf16 fc6 3 %, nothing for fp16x2; knob removed).
TAG=${1:-r02x}
mkdir -p gpurun_out
COMMON="--no-cpu-baseline --no-other-modes --no-e2e --no-verify --steps 100 --warmup 5 --precondition-s 0.3"
: > gpurun_out/${TAG}_ab.txt
for rep in 1 2; do for R in 0 1; do
  export SNN_DBG_PROD_REUSE=$R
  for CFG in "--workload bdd --batch 4 --mode bf16" ""; do
    timeout 300 python bench.py $CFG $COMMON > gpurun_out/${TAG}_one.json
    python - <<PY >> gpurun_out/${TAG}_ab.txt
import json
d=json.loads(open("gpurun_out/${TAG}_one.json").read().strip().splitlines()[-1]); c=d["config"]
print("reuse $R", c["workload"][:12], c["weight_mode"], "->", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if "fc" in k})
PY
  done
done; done
rm -f gpurun_out/${TAG}_one.json
sort gpurun_out/${TAG}_ab.txt
