#!/bin/bash
TAG=${1:-r01ao}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
COMMON="--steps 10 --warmup 3 --no-cpu-baseline --no-other-modes --no-e2e --mode fp16x2"
for V in "--t-rpn 32 --t-det 32" "--t-rpn 24 --t-det 24" "--t-rpn 8 --t-det 12"; do
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["launches_per_step"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_t.txt
