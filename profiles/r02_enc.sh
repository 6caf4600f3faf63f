#!/bin/bash
# Run on the GPU box: all GPU tests, the bench line, and ncu of the encoder / readout launches.  Usage: profiles/r02_enc.sh <tag>
TAG=${1:-r02j}
mkdir -p gpurun_out/parity
export SNN_PARITY_STATS_DIR=gpurun_out/parity
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), "e2e", round(d["e2e"]["value"],1), "verify", d["verify"]["ok"])
print({k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}); print({k:v for k,v in d["other_kernels"].items() if "encoder" in k})
PY
CMD="python bench.py --steps 2 --warmup 3 --precondition-s 0 --no-e2e --no-cpu-baseline --no-other-modes --no-verify"
ncu --set full --clock-control none -k regex:"encode_|readout_" -s 9 -c 3 -o gpurun_out/${TAG}_aux_fp16x2 $CMD > gpurun_out/${TAG}_aux.log 2>&1
ncu -i gpurun_out/${TAG}_aux_fp16x2.ncu-rep --page raw --csv > gpurun_out/${TAG}_aux_fp16x2_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_aux_fp16x2.ncu-rep
timeout 300 python profiles/bench_next_rows.py 2>&1 | tail -1 > gpurun_out/${TAG}_next_rows.json; cat gpurun_out/${TAG}_next_rows.json
