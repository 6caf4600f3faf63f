#!/bin/bash
# fc7 / fc6-tail ring depths (weight stages, spike-tile stages): SNN_DBG_STAGES applies where the rings fit (not the fc6 dual tiles)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-r02m}_stages.txt; : > $OUT
for S in "" "8,4" "7,5" "6,6" "5,7" "8,3"; do
  SNN_DBG_STAGES=$S timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-other-modes --no-verify > /tmp/b.json 2>/dev/null
  python - "$S" <<PY >> $OUT
import json, sys
d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
p=d["phase_ms_per_step"]
print("stages", repr(sys.argv[1]), "value", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), "fc6", round(p["fc6_lif_gemm"],4), "fc7", round(p["fc7_lif_gemm"],4), "conv", round(p["rpn_conv_lif_gemm"],4))
PY
done
cat $OUT
