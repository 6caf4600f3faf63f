#!/bin/bash
# Run on the GPU box (via gpurun): all GPU tests + bench of the three headline modes + conv role counters.
TAG=${1:-r01af}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 200 python scratch/time_roles.py 0 fp16x2 2>&1 | tail -7 | tee gpurun_out/${TAG}_roles.txt
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--mode fp16x2" "--mode bf16" "--mode fp32_exact"; do
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["launches_per_step"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_modes.txt
