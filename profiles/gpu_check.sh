#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests, then bench lines for the weight modes.
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
for MODE in fp32_exact fp16x2 bf16 fp16; do
  timeout 600 python bench.py --mode $MODE --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${MODE}.json 2> gpurun_out/${TAG}_bench_${MODE}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${MODE}.json").read().strip().splitlines()[-1])
    print("${MODE}", round(d["value"],1), "img/s", "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v}, "roof", round(d["roofline"]["achieved"],1), d["clocks"])
except Exception as e:
    print("${MODE} failed", e); print(open("gpurun_out/${TAG}_bench_${MODE}.err").read()[-2000:])
PY
done
