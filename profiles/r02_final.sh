#!/bin/bash
TAG=${1:-r02r}
bash profiles/r02_check.sh $TAG
COMMON="--no-cpu-baseline --no-other-modes --no-verify --no-e2e --steps 100 --warmup 5"
: > gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --t-rpn 4 --t-det 4 $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --t-rpn 12 --t-det 16 $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 $COMMON >> gpurun_out/${TAG}_configs.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(l); c=d["config"]
    print(c["workload"][:12], c["weight_mode"], "T", c["T_rpn"], c["T_det"], "->", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v})
PY
