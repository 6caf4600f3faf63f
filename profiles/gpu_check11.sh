#!/bin/bash
TAG=${1:-r01ar}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
COMMON="--steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--mode fp16x2" "--mode bf16" "--workload bdd --batch 4 --mode bf16" "--mode fp16x2 --t-rpn 24 --t-det 24"; do
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}, d["roofline"].get("frac_of_effective_clock_ceiling"), d["roofline"].get("in_kernel",{}).get("effective_sm_mhz"))
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_modes.txt
timeout 200 python scratch/time_roles.py 1 bf16 2>&1 | tail -13 | head -6
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dual_tiles_currents or encoder_rows_bit_exact" > gpurun_out/${TAG}_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/${TAG}_sanitizer.log
tail -6 gpurun_out/${TAG}_sanitizer.log
