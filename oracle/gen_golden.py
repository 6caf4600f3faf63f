"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz.

Runs the UNMODIFIED reference modules (/root/reference/rpn.py::RPNHeadSNN,
/root/reference/faster_rcnn.py::FastRCNNPredictorSNNFull) on CPU with
oracle/norse_shim standing in for the missing norse==0.0.7 package, and records
their outputs plus the per-step spikes of every LIF layer (captured with
forward hooks, so no reference source is edited or copied).

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py
The GPU box never runs this; tests read the committed .npz files.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "norse_shim"))
    sys.path.insert(0, REF)
    import rpn  # noqa: E402
    import faster_rcnn  # noqa: E402
    return rpn, faster_rcnn


def _pack(spk_list):
    """list of T tensors {0,1} -> uint32 spike-train words (bit t = step t)."""
    w = torch.zeros(spk_list[0].shape, dtype=torch.int64)
    for t, s in enumerate(spk_list):
        w |= s.to(torch.int64) << t
    return w.numpy().astype(np.uint32)


def _hook(store):
    def fn(_mod, _inp, out):
        store.append(out[0].detach().clone())
    return fn


def gen_rpn(rpn_mod, name, in_channels, num_anchors, T, levels, N, seed, scale, store_tensors):
    torch.manual_seed(seed)
    head = rpn_mod.RPNHeadSNN(in_channels, num_anchors, T)
    g = torch.Generator().manual_seed(seed + 1000)
    feats = [scale * torch.randn(N, in_channels, h, w, generator=g) for (h, w) in levels]
    rec = []
    head.shared_lif.register_forward_hook(_hook(rec))
    with torch.no_grad():
        logits, bbox = head(feats)
    d = {"kind": "rpn", "in_channels": in_channels, "num_anchors": num_anchors, "T": T, "N": N,
         "levels": np.array(levels), "seed": seed, "scale": scale,
         "w_checksum": np.array([float(p.detach().double().sum()) for p in
                                 (head.shared_conv.weight, head.conv_cls.weight, head.conv_bbox.weight)])}
    for l in range(len(levels)):
        d[f"logits{l}"] = logits[l].numpy()
        d[f"bbox{l}"] = bbox[l].numpy()
        d[f"trains{l}"] = _pack(rec[l * T:(l + 1) * T])            # [N,C,H,W] uint32
    if store_tensors:
        d["w_shared"] = head.shared_conv.weight.detach().numpy()
        d["w_cls"] = head.conv_cls.weight.detach().numpy()
        d["w_bbox"] = head.conv_bbox.weight.detach().numpy()
        for l, f in enumerate(feats):
            d[f"feat{l}"] = f.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    rate = np.mean([np.mean([float(r.mean()) for r in rec])])
    print(f"{name}: shared-LIF mean rate/step {rate:.4f}, |logit|max {max(float(x.abs().max()) for x in logits):.4f}")


def gen_box(frcnn_mod, name, in_shape, rep, C, T, R, seed, scale, only_one_bbox, store_tensors):
    in_channels = int(np.prod(in_shape))
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        head = frcnn_mod.FastRCNNPredictorSNNFull(in_channels, rep, C, T, only_one_bbox=only_one_bbox)
    g = torch.Generator().manual_seed(seed + 1000)
    x = scale * torch.randn(R, *in_shape, generator=g)
    r6, r7 = [], []
    head.lif6.register_forward_hook(_hook(r6))
    head.lif7.register_forward_hook(_hook(r7))
    with torch.no_grad():
        cls, box = head(x)
    d = {"kind": "box", "in_shape": np.array(in_shape), "rep": rep, "C": C, "T": T, "R": R, "seed": seed,
         "scale": scale, "only_one_bbox": only_one_bbox,
         "cls": cls.numpy(), "bbox": box.numpy(), "trains6": _pack(r6), "trains7": _pack(r7),
         "w_checksum": np.array([float(p.detach().double().sum()) for p in
                                 (head.fc6.weight, head.fc7.weight, head.cls_score.weight, head.bbox_pred.weight)])}
    if store_tensors:
        d.update(w6=head.fc6.weight.detach().numpy(), w7=head.fc7.weight.detach().numpy(),
                 w_cls=head.cls_score.weight.detach().numpy(), w_bbox=head.bbox_pred.weight.detach().numpy(),
                 x=x.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: spk6 rate {np.mean([float(r.mean()) for r in r6]):.4f} spk7 rate "
          f"{np.mean([float(r.mean()) for r in r7]):.4f} |cls|max {float(cls.abs().max()):.4f}")


def main():
    torch.set_num_threads(1)          # fixed accumulation blocking for the goldens
    os.makedirs(OUT, exist_ok=True)
    rpn_mod, frcnn_mod = _import_reference()
    # tiny cases: every tensor stored, so the oracle tests need no RNG reproduction
    gen_rpn(rpn_mod, "rpn_tiny", 16, 3, 6, [(5, 7), (3, 4)], 2, seed=11, scale=12.0, store_tensors=True)
    gen_box(frcnn_mod, "box_tiny", (3, 4, 4), 32, 5, 7, 9, seed=12, scale=6.0, only_one_bbox=False, store_tensors=True)
    gen_box(frcnn_mod, "box_tiny_onebbox", (3, 4, 4), 32, 5, 7, 9, seed=13, scale=6.0, only_one_bbox=True,
            store_tensors=True)
    # kernel-shaped cases: inputs/weights re-drawn from the seeds at test time
    gen_rpn(rpn_mod, "rpn_c256_T8", 256, 3, 8, [(12, 24), (6, 11)], 2, seed=21, scale=1.0, store_tensors=False)
    gen_rpn(rpn_mod, "rpn_c256_T12", 256, 3, 12, [(9, 16)], 1, seed=22, scale=1.0, store_tensors=False)
    gen_box(frcnn_mod, "box_k12544_T12", (256, 7, 7), 1024, 9, 12, 40, seed=23, scale=1.0, only_one_bbox=False,
            store_tensors=False)
    gen_box(frcnn_mod, "box_k12544_T8_onebbox", (256, 7, 7), 1024, 5, 8, 24, seed=24, scale=1.0, only_one_bbox=True,
            store_tensors=False)


if __name__ == "__main__":
    main()
