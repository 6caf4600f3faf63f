#!/bin/bash
# Run on the GPU box (via gpurun): kernel tests, fc6 tail variants, ncu of the two fc6 launches.
TAG=${1:-r01ae}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/${TAG}_pytest_kernels.log 2>&1
echo "pytest kernels rc=$?" >> gpurun_out/${TAG}_pytest_kernels.log
tail -4 gpurun_out/${TAG}_pytest_kernels.log
COMMON="--mode fp16x2 --steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for V in "--fc-no-split 0" "--fc-no-split 2" "--fc-no-split 1" "--fc-no-split 0 --mode bf16" "--fc-no-split 0 --mode fp32_exact"; do
  timeout 300 python bench.py $COMMON $V > gpurun_out/${TAG}_bench_v.json 2> gpurun_out/${TAG}_bench_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_v.json").read().strip().splitlines()[-1])
    print("$V", round(d["value"],1), "img/s", {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if 'fc' in k}, d["launches_per_step"])
except Exception as e:
    print("$V failed", e); print(open("gpurun_out/${TAG}_bench_v.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_tail.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 13 -c 2 -o gpurun_out/${TAG}_fc6 \
    python bench.py --steps 2 --warmup 3 --mode fp16x2 --no-e2e --no-cpu-baseline --no-other-modes > gpurun_out/${TAG}_fc6.log 2>&1
ls -la gpurun_out/${TAG}_fc6.ncu-rep
