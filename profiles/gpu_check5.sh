#!/bin/bash
# Run on the GPU box (via gpurun): where the MMA thread waits (fc6 dual, conv), fc ring-depth sweep, ncu of fc6.
TAG=${1:-r01ad}
mkdir -p gpurun_out
for PH in 1 0 2; do timeout 200 python scratch/time_roles.py $PH fp16x2 2>&1 | tail -8; done | tee gpurun_out/${TAG}_roles.txt
COMMON="--mode fp16x2 --steps 20 --warmup 5 --no-cpu-baseline --no-other-modes --no-e2e"
for ST in "6,4" "5,5" "4,5" "3,6" "2,7"; do
  SNN_DBG_STAGES=$ST timeout 300 python bench.py $COMMON > gpurun_out/${TAG}_bench_st.json 2> gpurun_out/${TAG}_bench_st.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_st.json").read().strip().splitlines()[-1])
    print("stages $ST", round(d["value"],1), "img/s", {k: round(v,4) for k,v in d["phase_ms_per_step"].items() if 'fc' in k})
except Exception as e:
    print("$ST failed", e); print(open("gpurun_out/${TAG}_bench_st.err").read()[-1500:])
PY
done | tee gpurun_out/${TAG}_stages.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 10 -c 1 -o gpurun_out/${TAG}_fc6 \
    python bench.py --steps 2 --warmup 3 --mode fp16x2 --no-e2e --no-cpu-baseline --no-other-modes > gpurun_out/${TAG}_fc6.log 2>&1
ls -la gpurun_out/${TAG}_fc6.ncu-rep
