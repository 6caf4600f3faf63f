#!/usr/bin/env python
"""(T_rpn, T_det) sweep of the reference's energy estimate, in the shape of its own sweep file
(metrics_for_different_timesteps.py:360-377, 495-508 with `-o efficiency`: a JSON list of [t_rpn, t_det, value]).

The reference rebuilds the model and reloads the checkpoint for every pair; here T is a run-time argument of the kernels
and the weights are T-independent, so one pair of modules serves the whole sweep.  The value is train.py:470-517's
SNN / ANN energy ratio (spikes x FLOPs x 0.9 pJ over FLOPs x 4.6 pJ, layers shared_lif per FPN level + lif6 + lif7)
computed from the spike trains the kernels emit (rates.py).  Synthetic Cityscapes-shaped input, random-init weights.
Usage (GPU box): python profiles/energy_sweep.py [--rpn1 4 --rpn2 12 --det1 8 --det2 16] [--out FILE]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snn_automotive_object_detection_b200 as S
from bench import bench_inputs, WORKLOADS, CH, HID, KBOX


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rpn1", type=int, default=4); ap.add_argument("--rpn2", type=int, default=12)      # the reference's
    ap.add_argument("--det1", type=int, default=8); ap.add_argument("--det2", type=int, default=16)      # default ranges
    ap.add_argument("--workload", default="cityscapes")
    ap.add_argument("--mode", default="fp16x2")
    ap.add_argument("--out", default="gpurun_out/energy_sweep.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    C = WORKLOADS[a.workload]["classes"]
    torch.manual_seed(0)
    rpn = S.RPNHeadSNN(CH, 3, a.rpn1, mode=a.mode).to(dev).eval()
    box = S.FastRCNNPredictorSNNFull(KBOX, HID, C, a.det1, mode=a.mode).to(dev).eval()
    rpn.record_spikes = box.record_spikes = True
    rpn.record_rates = box.record_rates = True
    f, r = bench_inputs(a.workload, 0)
    feats = [x.unsqueeze(0).to(dev) for x in f]
    rois = r.to(dev)
    results, detail = [], []
    t0 = time.perf_counter()
    for t_rpn in range(a.rpn1, a.rpn2 + 1):
        rpn.num_steps = t_rpn
        rpn(feats)
        rr = S.rpn_spike_rates_and_flops(rpn)
        for t_det in range(a.det1, a.det2 + 1):
            box.num_steps = t_det
            box(rois)
            br = S.box_spike_rates_and_flops(box)
            ratio, layers = S.energy_report(rr, br, t_rpn, t_det)
            results.append([t_rpn, t_det, ratio])
            detail.append({"t_rpn": t_rpn, "t_det": t_det, "layers": layers})
    torch.cuda.synchronize()
    with open(a.out, "w") as fp:
        json.dump(results, fp)
    with open(a.out.replace(".json", "_layers.json"), "w") as fp:
        json.dump(detail, fp)
    print(f"{len(results)} (T_rpn, T_det) pairs in {time.perf_counter() - t0:.1f} s -> {a.out}")
    print(results[:3], "...", results[-1])


if __name__ == "__main__":
    main()
