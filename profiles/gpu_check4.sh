#!/bin/bash
# Run on the GPU box (via gpurun): all GPU tests, headline bench, ncu --set full of the two encoders.
TAG=${1:-r01ac}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
COMMON="--mode fp16x2 --steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--fc-dual 0"; do
  N=$(echo $V | tr -d ' -')
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_${N}.json 2> gpurun_out/${TAG}_bench_${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${N}.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["clocks"]["sm_mhz"], d["launches_per_step"])
    print(d["other_kernels"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_${N}.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_ -s 6 -c 2 -o gpurun_out/${TAG}_enc \
    python bench.py --steps 2 --warmup 3 --mode fp16x2 --no-e2e --no-cpu-baseline --no-other-modes > gpurun_out/${TAG}_enc.log 2>&1
ls -la gpurun_out/${TAG}_enc.ncu-rep
