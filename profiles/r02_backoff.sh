#!/bin/bash
# Conv producers: sleep between polls of the stage they wait for.  A/B over SNN_DBG_BACKOFF (ns).  Usage: profiles/r02_backoff.sh <tag>
TAG=${1:-r02q}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_backoff.txt; : > $OUT
for NS in 0 100 300 1000 0 300 3000; do
  SNN_DBG_BACKOFF=$NS timeout 300 python bench.py --steps 400 --warmup 5 --no-e2e --no-cpu-baseline --no-other-modes > /tmp/b.json 2>/dev/null
  python - $NS <<PY >> $OUT
import json, sys
d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
p=d["phase_ms_per_step"]; r=d["roofline"]
print("backoff_ns", sys.argv[1], "value", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), "conv ms", round(p["rpn_conv_lif_gemm"],4), "in-kernel MHz", round(r["in_kernel"]["effective_sm_mhz"],1), "frac_eff", round(r["frac_of_effective_clock_ceiling"],3), "power", round(d["clocks"]["power_w_nvml_trailing_avg"],1), "verify", d["verify"] and d["verify"]["ok"])
PY
done
cat $OUT
SNN_DBG_BACKOFF=300 timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-other-modes --no-verify > /tmp/b.json 2>/dev/null
python -c "
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]); print('bdd bf16 backoff 300:', round(d['value'],1), round(d['phase_ms_per_step']['rpn_conv_lif_gemm'],4))" | tee -a $OUT
