#!/bin/bash
# Run on a 2-GPU box (gpurun --gpus 2): the driver's own N = 2 launch of both bench arms on the final code.
TAG=${1:-r02n2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 0 > gpurun_out/${TAG}_reference_2.json 2> gpurun_out/${TAG}_reference_2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_2.json 2> gpurun_out/${TAG}_bench_2.err
python - <<PY
import json
for f in ("reference_2", "bench_2"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], "n_gpus", d["n_gpus"], "e2e", round(d["e2e"]["value"],1), d.get("clocks"), d.get("verify") and d["verify"]["ok"], d.get("ms_per_step_by_rank"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/${TAG}_bench_2.err
