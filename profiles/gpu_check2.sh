#!/bin/bash
# Run on the GPU box (via gpurun): the new GPU tests first (bounded), then short bench lines of the fc tiling variants.
TAG=${1:-r01aa}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/${TAG}_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/${TAG}_pytest_kernels.log
tail -15 gpurun_out/${TAG}_pytest_kernels.log
COMMON="--mode fp16x2 --steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--fc-dual 1" "--fc-dual 0" "--fc-dual 1 --fc-units 16" "--fc-dual 0 --fc-units 16"; do
  N=$(echo $V | tr -d ' -')
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_${N}.json 2> gpurun_out/${TAG}_bench_${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${N}.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["clocks"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_${N}.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_kernels.py > gpurun_out/${TAG}_pytest_rest.log 2>&1
echo "pytest rest rc=$?" >> gpurun_out/${TAG}_pytest_rest.log
tail -8 gpurun_out/${TAG}_pytest_rest.log
