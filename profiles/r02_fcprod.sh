#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_heads.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
COMMON="--no-cpu-baseline --no-other-modes --no-e2e --steps 100 --warmup 5"
: > gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 $COMMON --no-verify >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --mode bf16 $COMMON --no-verify >> gpurun_out/${TAG}_configs.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(l); c=d["config"]
    print(c["workload"][:12], c["weight_mode"], "->", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v}, d["verify"] and d["verify"]["ok"])
PY
