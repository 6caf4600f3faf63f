#!/bin/bash
# Conv weight multicast experiment: bit-identity test, then A/B of the bench (sustained + 1200 steps + NVML power) and the
# L2 / crossbar counters of the conv launch under ncu.  Usage: profiles/r02_mc.sh <tag>
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_heads.py -m gpu -q -x -k "multicast" > gpurun_out/${TAG}_mc_test.log 2>&1; echo "mc test rc=$?"; tail -5 gpurun_out/${TAG}_mc_test.log
grep -q "passed" gpurun_out/${TAG}_mc_test.log || exit 0
COMMON="--no-e2e --no-cpu-baseline --no-other-modes --warmup 5"
for MC in 0 1 0 1; do
  timeout 300 python bench.py $COMMON --steps 400 --conv-multicast $MC > /tmp/b.json 2>/tmp/b.err
  python - $MC <<PY | tee -a gpurun_out/${TAG}_mc_ab.txt
import json, sys
d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
p=d["phase_ms_per_step"]; r=d["roofline"]
print("multicast", sys.argv[1], "value", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), "conv ms", round(p["rpn_conv_lif_gemm"],4), "in-kernel MHz", round(r["in_kernel"]["effective_sm_mhz"],1), "frac_eff", round(r["frac_of_effective_clock_ceiling"],3), "power", d["clocks"]["power_w_nvml_trailing_avg"], "sm", d["clocks"]["sm_mhz"], "verify", d["verify"] and d["verify"]["ok"])
PY
done
timeout 300 python bench.py $COMMON --no-verify --steps 1200 --conv-multicast 1 > gpurun_out/${TAG}_bench_1200_mc.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_1200_mc.json').read().strip().splitlines()[-1]); print('1200 steps multicast:', round(d['value'],1), d['clocks'])" | tee -a gpurun_out/${TAG}_mc_ab.txt
CMD="python bench.py --steps 2 --warmup 3 --precondition-s 0 --no-e2e --no-cpu-baseline --no-other-modes --no-verify --conv-multicast 1"
ncu --set full --clock-control none -k regex:spike_gemm_lif -s 12 -c 1 -o gpurun_out/${TAG}_conv_mc $CMD > gpurun_out/${TAG}_conv_mc.log 2>&1
ncu -i gpurun_out/${TAG}_conv_mc.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_mc_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_conv_mc.ncu-rep
