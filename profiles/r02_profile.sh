#!/bin/bash
# Run on the GPU box (via gpurun, 1 GPU): the evidence set of round 2.  Usage: profiles/r02_profile.sh <tag>
TAG=${1:-r02b}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --precondition-s 0 --no-e2e --no-cpu-baseline --no-other-modes --no-verify"
# launch list of one short run (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_fp16x2.csv $CMD --mode fp16x2 > gpurun_out/${TAG}_launches.log 2>&1
# --set full of the spike GEMM launches of one step (conv, fc6 dual, fc6 tail, fc7) and of the aux kernels, headline mode
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 12 -c 4 -o gpurun_out/${TAG}_gemm_fp16x2 $CMD --mode fp16x2 > gpurun_out/${TAG}_gemm_fp16x2.log 2>&1
ncu --set full --clock-control none -k regex:"encode_|readout_" -s 9 -c 3 -o gpurun_out/${TAG}_aux_fp16x2 $CMD --mode fp16x2 > gpurun_out/${TAG}_aux_fp16x2.log 2>&1
# BASELINE config 3: BDD batch 4, 5 classes, bf16 -- the conv + fc6 launches with source
ncu --set full --clock-control none --import-source on -k regex:spike_gemm_lif -s 12 -c 4 -o gpurun_out/${TAG}_gemm_bdd_bf16 $CMD --mode bf16 --workload bdd --batch 4 > gpurun_out/${TAG}_gemm_bdd_bf16.log 2>&1
# fused RoIAlign kernels
ncu --set full --clock-control none -k regex:roi_align -s 6 -c 2 -o gpurun_out/${TAG}_roi python profiles/bench_next_rows.py > gpurun_out/${TAG}_roi.log 2>&1
# gpurun brings back at most 64 MiB: export the pages read here (raw metrics; per-line source counters of the GEMMs) on
# the box and drop the .ncu-rep files
for R in gemm_fp16x2 aux_fp16x2 gemm_bdd_bf16 roi; do
  F=gpurun_out/${TAG}_${R}.ncu-rep
  [ -f $F ] || continue
  ncu -i $F --page raw --csv > gpurun_out/${TAG}_${R}_raw.csv 2>/dev/null
  case $R in gemm_*) ncu -i $F --page source --csv --print-source sass > gpurun_out/${TAG}_${R}_source.csv 2>/dev/null; gzip -f gpurun_out/${TAG}_${R}_source.csv;; esac
  rm -f $F
done
ls -la gpurun_out | grep ${TAG}
# timed (no profiler): next rows, configs 3 / 4 / 5 (N = 1 leg), energy sweep, long run with NVML power
timeout 300 python profiles/bench_next_rows.py > gpurun_out/${TAG}_next_rows.json 2> gpurun_out/${TAG}_next_rows.err; cat gpurun_out/${TAG}_next_rows.json
COMMON="--no-cpu-baseline --no-other-modes --no-verify --steps 100 --warmup 5"
: > gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --workload bdd --batch 4 --mode bf16 $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --workload bdd --batch 4 --mode fp16x2 $COMMON >> gpurun_out/${TAG}_configs.jsonl
timeout 300 python bench.py --global-batch 64 $COMMON --no-e2e --steps 20 >> gpurun_out/${TAG}_configs.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(l); c=d["config"]
    print(c["workload"][:12], c["weight_mode"], "B", c["global_batch"], "->", round(d["value"],1), "img/s burst", d["first_20_steps"] and round(d["first_20_steps"]["value"],1), "ms/step", round(d["ms_per_step"],3), "roof", round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_of_effective_clock_ceiling",0),3), {k: round(v,3) for k,v in d["phase_ms_per_step"].items() if v})
PY
timeout 600 python profiles/t_sweep.py --out gpurun_out/${TAG}_t_sweep.json 2>&1 | tail -3
timeout 600 python profiles/energy_sweep.py --out gpurun_out/${TAG}_energy_sweep.json 2>&1 | tail -3
timeout 300 python bench.py --steps 1200 --no-cpu-baseline --no-other-modes --no-verify --no-e2e > gpurun_out/${TAG}_bench_1200_steps.json 2> gpurun_out/${TAG}_bench_1200.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_1200_steps.json").read().strip().splitlines()[-1])
print("1200 steps:", round(d["value"],1), "img/s", d["clocks"])
PY
