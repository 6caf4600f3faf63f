#!/bin/bash
TAG=${1:-r01av}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for V in 1 0 1 0; do
  SNN_PDL=$V timeout 300 python bench.py --no-cpu-baseline --no-other-modes --no-e2e > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_b.json").read().strip().splitlines()[-1])
    print("PDL $V", round(d["value"],1), round(d["ms_per_step"],4), "first20", round(d["first_20_steps"]["value"],1), d["roofline"]["in_kernel"]["effective_sm_mhz"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_b.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_pdl.txt
