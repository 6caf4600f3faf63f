#!/bin/bash
# Run on the GPU box (via gpurun): the measurement set of a round.  Usage: profiles/final_round.sh <tag>
TAG=${1:-r01ah}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 600 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print("default", round(d["value"],1), "img/s e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1), d["roofline"].get("frac_of_clock_ceiling"), d["clocks"], d["cpu_baseline"]["value"], d["other_modes"])
print({k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}); print(d["other_kernels"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 300 gpurun_out/${TAG}_bench_reference.json
bash profiles/run_ncu.sh ${TAG} fp16x2 > gpurun_out/${TAG}_run_ncu.log 2>&1
bash profiles/run_configs.sh ${TAG} 2>&1 | tail -12
for M in bf16 fp16x2; do timeout 200 python scratch/time_roles.py 0 $M 2>&1 | tail -9; timeout 200 python scratch/time_roles.py 1 $M 2>&1 | tail -9; done | tee gpurun_out/${TAG}_roles.txt
