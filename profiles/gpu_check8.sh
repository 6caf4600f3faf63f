#!/bin/bash
TAG=${1:-r01ag}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e --mode fp16x2"
for V in "A" "SNN_BENCH_NO_PHASES=1" "A" "SNN_BENCH_NO_PHASES=1"; do
  env $V timeout 300 python bench.py $COMMON > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["launches_per_step"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_phases.txt
