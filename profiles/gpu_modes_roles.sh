#!/bin/bash
TAG=${1:-r01ai}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_heads.py -m gpu -x -q 2>&1 | tail -3
for M in bf16 fp16x2; do timeout 200 python scratch/time_roles.py 0 $M 2>&1 | tail -9; done | tee gpurun_out/${TAG}_roles.txt
COMMON="--steps 100 --warmup 10 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--mode fp16x2" "--mode bf16" "--mode fp32_exact"; do
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v})
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_modes.txt
