COMMON="--no-cpu-baseline --no-other-modes --no-e2e --no-verify --steps 100 --warmup 5 --precondition-s 0.3"
for rep in 1 2; do for S in 4,5 3,6 5,4 6,4; do
  export SNN_DBG_STAGES=$S
  timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 $COMMON > /tmp/one.json
  python - <<PY
import json
d=json.loads(open("/tmp/one.json").read().strip().splitlines()[-1])
print("stages $S ->", round(d["value"],1), "burst", round(d["first_20_steps"]["value"],1), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if "fc" in k})
PY
done; done
