#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): BASELINE config 5 -- 64 Cityscapes-shaped images per step, image-sharded over
# 1 / 2 / 4 / 8 B200s (strong scaling; reference sharding: DistributedSampler + DDP, train.py:594-601,708), NCCL gather of
# the per-image spike-rate records.  Usage: profiles/r02_scale.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
COMMON="--global-batch 64 --steps 20 --warmup 3 --no-cpu-baseline --no-other-modes --no-verify --no-e2e"
timeout 300 python bench.py --gpus 1 $COMMON > gpurun_out/${TAG}_config5_1.json 2> gpurun_out/${TAG}_config5_1.err
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N $COMMON > gpurun_out/${TAG}_config5_${N}.json 2> gpurun_out/${TAG}_config5_${N}.err
done
# weak scaling with the end-to-end leg at N = 8 (host-side limits), default 2 images per GPU
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
    bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --no-other-modes > gpurun_out/${TAG}_scale_8.json 2> gpurun_out/${TAG}_scale_8.err
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_config5_{n}.json").read().strip().splitlines()[-1])
        base = base or d["value"]
        print("config5 N", n, "img/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "eff", round(d["value"]/(n*base),3), "burst", d["first_20_steps"] and round(d["first_20_steps"]["value"],1), [round(x,2) for x in d["ms_per_step_by_rank"]])
    except Exception as e:
        print("config5 N", n, "failed", e)
try:
    d=json.loads(open("gpurun_out/${TAG}_scale_8.json").read().strip().splitlines()[-1])
    print("weak N 8:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "e2e_fused", d["e2e_fused_roi_pool"], d["host"])
except Exception as e:
    print("weak 8 failed", e)
PY
