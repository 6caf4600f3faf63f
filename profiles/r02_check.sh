#!/bin/bash
# Run on the GPU box (via gpurun): tests + smoke + the bench line.  Usage: profiles/r02_check.sh <tag> [pytest-args]
TAG=${1:-r02a}
mkdir -p gpurun_out/parity
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt
export SNN_PARITY_STATS_DIR=gpurun_out/parity
timeout 1500 python -m pytest tests -m gpu -x -q ${2} > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -5 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench rc=$?"; tail -c 800 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("value", round(d["value"],1), "burst", d["first_20_steps"] and round(d["first_20_steps"]["value"],1), "e2e", round(d["e2e"]["value"],1), "e2e_fused", d["e2e_fused_roi_pool"])
    print("roof", {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ("achieved","frac","frac_of_burst_peak","frac_of_effective_clock_ceiling","ms_per_launch","regime")}, r.get("in_kernel",{}).get("effective_sm_mhz"))
    print("clocks", d["clocks"]); print("host", d["host"]); print("canonical", d["canonical"])
    print("cpu", d["cpu_baseline"]); print("verify", d["verify"])
    print({k: round(v,4) for k,v in d["phase_ms_per_step"].items() if v}); print(d["other_kernels"]); print(d["other_modes"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 400 gpurun_out/${TAG}_bench_reference.json
