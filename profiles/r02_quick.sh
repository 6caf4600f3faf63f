#!/bin/bash
# Run on the GPU box: tests + the bench line + config 3 + role counters.  Usage: profiles/r02_quick.sh <tag>
TAG=${1:-r02d}
mkdir -p gpurun_out/parity
export SNN_PARITY_STATS_DIR=gpurun_out/parity
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
COMMON="--no-cpu-baseline --no-other-modes --steps 100 --warmup 5"
: > gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --mode bf16 $COMMON --no-e2e >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --t-rpn 4 --t-det 4 $COMMON --no-e2e >> gpurun_out/${TAG}_configs.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(l); c=d["config"]
    print(c["workload"][:12], c["weight_mode"], "B", c["global_batch"], "T", c["T_rpn"], c["T_det"], "->", round(d["value"],1), "img/s burst", d["first_20_steps"] and round(d["first_20_steps"]["value"],1), "ms/step", round(d["ms_per_step"],3), "roof", round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_of_effective_clock_ceiling",0),3), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v}, "verify", d["verify"] and d["verify"]["ok"])
PY
for M in bf16 fp16x2; do timeout 200 python scratch/time_roles.py 0 $M 2>&1 | tail -12; done | tee gpurun_out/${TAG}_roles.txt
