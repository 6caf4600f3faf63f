#!/usr/bin/env python
"""Summarise .ncu-rep captures (brought back in gpurun_out/) into small CSVs under profiles/.
Usage: python profiles/summarise_ncu.py gpurun_out/<tag>_gemm_<mode>.ncu-rep | gpurun_out/<tag>_..._raw.csv [...]
(a *_raw.csv is the `ncu -i <rep> --page raw --csv` export made on the GPU box: gpurun returns at most 64 MiB)"""
import csv
import io
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for rep in sys.argv[1:]:
        if rep.endswith(".csv"):
            out = open(rep).read()
        else:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        name = os.path.splitext(os.path.basename(rep))[0].replace("_raw", "") + "_summary.csv"
        with open(os.path.join(here, name), "w", newline="") as f:
            w = csv.writer(f)
            # every tensor-pipe counter the capture holds (the guide's sm__pipe_tensor_cycles_active among them)
            tensor = [h for h in hdr if h.startswith("sm__pipe_tensor") or h.startswith("sm__inst_executed_pipe_tensor")]
            keep = ["Kernel Name"] + [m for m in METRICS if m in col] + [h for h in tensor if h not in METRICS]
            w.writerow(keep)
            w.writerow([units[col[k]] for k in keep])
            for r in rows[2:]:
                w.writerow([r[col[k]] for k in keep])
        print("wrote", name)


if __name__ == "__main__":
    main()
