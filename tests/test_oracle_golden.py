"""The oracle port (oracle/snn_oracle.py) against golden vectors produced by the
UNMODIFIED reference modules (oracle/gen_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import snn_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _rpn_inputs(g):
    """Re-draw weights/features exactly as oracle/gen_golden.py did (seed-only fixtures)."""
    C, A, T, N = int(g["in_channels"]), int(g["num_anchors"]), int(g["T"]), int(g["N"])
    if "w_shared" in g:
        w = [torch.from_numpy(g[k]) for k in ("w_shared", "w_cls", "w_bbox")]
        feats = [torch.from_numpy(g[f"feat{l}"]) for l in range(len(g["levels"]))]
    else:
        W = O.reference_weights(in_channels=C, num_anchors=A, box_in=8, rep=8, num_classes=2, seed=int(g["seed"]))
        w = [W["shared_conv"], W["conv_cls"], W["conv_bbox"]]
        gen = torch.Generator().manual_seed(int(g["seed"]) + 1000)
        feats = [float(g["scale"]) * torch.randn(N, C, int(h), int(wd), generator=gen) for h, wd in g["levels"]]
    chk = np.array([float(x.double().sum()) for x in w])
    assert np.allclose(chk, g["w_checksum"], rtol=0, atol=1e-9), "weight reproduction drifted"
    return w, feats, T


@pytest.mark.parametrize("name", ["rpn_tiny", "rpn_c256_T8", "rpn_c256_T12"])
def test_rpn_port_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    w, feats, T = _rpn_inputs(g)
    torch.set_num_threads(1)
    lo, bb, tr = O.rpn_head_forward(feats, *w, T, record=True)
    for l in range(len(feats)):
        trains = O.pack_trains(tr[l]["spk"]).numpy().astype(np.uint32)
        agree = (trains == g[f"trains{l}"]).mean()
        assert agree == 1.0, f"level {l}: spike-train agreement {agree}"
        assert np.allclose(lo[l].numpy(), g[f"logits{l}"], rtol=0, atol=1e-6)
        assert np.allclose(bb[l].numpy(), g[f"bbox{l}"], rtol=0, atol=1e-6)
    assert sum(int(g[f"trains{l}"].astype(bool).sum()) for l in range(len(feats))) > 0


def _box_inputs(g):
    T, R, C, rep = int(g["T"]), int(g["R"]), int(g["C"]), int(g["rep"])
    in_shape = tuple(int(v) for v in g["in_shape"])
    if "w6" in g:
        w = [torch.from_numpy(g[k]) for k in ("w6", "w7", "w_cls", "w_bbox")]
        x = torch.from_numpy(g["x"])
    else:
        torch.manual_seed(int(g["seed"]))
        fc6 = torch.nn.Linear(int(np.prod(in_shape)), rep, bias=False)
        fc7 = torch.nn.Linear(rep, rep, bias=False)
        cs = torch.nn.Linear(rep, C, bias=False)
        bp = torch.nn.Linear(rep, 4 if bool(g["only_one_bbox"]) else 4 * C, bias=False)
        w = [m.weight.detach() for m in (fc6, fc7, cs, bp)]
        gen = torch.Generator().manual_seed(int(g["seed"]) + 1000)
        x = float(g["scale"]) * torch.randn(R, *in_shape, generator=gen)
    chk = np.array([float(t.double().sum()) for t in w])
    assert np.allclose(chk, g["w_checksum"], rtol=0, atol=1e-9), "weight reproduction drifted"
    return w, x, T


@pytest.mark.parametrize("name", ["box_tiny", "box_tiny_onebbox", "box_k12544_T12", "box_k12544_T8_onebbox"])
def test_box_port_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    w, x, T = _box_inputs(g)
    torch.set_num_threads(1)
    cls, box, tr = O.box_head_forward(x, *w, T, record=True)
    assert (O.pack_trains(tr["spk6"]).numpy().astype(np.uint32) == g["trains6"]).all()
    assert (O.pack_trains(tr["spk7"]).numpy().astype(np.uint32) == g["trains7"]).all()
    assert np.allclose(cls.numpy(), g["cls"], rtol=0, atol=1e-6)
    assert np.allclose(box.numpy(), g["bbox"], rtol=0, atol=1e-6)
    assert box.shape[1] == (4 if bool(g["only_one_bbox"]) else 4 * int(g["C"]))
