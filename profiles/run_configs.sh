#!/bin/bash
# Run on the GPU box: the BASELINE.json configs that are not the headline bench line.
#   config 3: BDD-shaped batch 4, 5 classes, bf16 weights        config 4: timestep sweep T in {4,8,12,16,32}
#   config 5 (N=1 leg): 64 Cityscapes-shaped images in one step  (the multi-GPU legs run under torchrun)
TAG=${1:-r01}
OUT=gpurun_out/${TAG}_configs.jsonl
: > $OUT
COMMON="--no-cpu-baseline --no-other-modes --steps 10 --warmup 3"
python bench.py --workload bdd --batch 4 --mode bf16 $COMMON >> $OUT
python bench.py --workload bdd --batch 4 --mode fp16x2 $COMMON >> $OUT
for T in 4 8 12 16 24 32; do
  python bench.py --t-rpn $T --t-det $T $COMMON --no-e2e >> $OUT
done
python bench.py --global-batch 64 $COMMON --no-e2e >> $OUT
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); c=d["config"]
    print(c["workload"][:12], c["weight_mode"], "B", c["global_batch"], "T", c["T_rpn"], c["T_det"], "->", round(d["value"],1), "img/s", "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v})
PY
