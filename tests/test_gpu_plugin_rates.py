"""Spike-rate report (reference section a8 format), the torchvision plug-in point, full-size parity."""
import os

import numpy as np
import pytest
import torch

from oracle import snn_oracle as O
import snn_automotive_object_detection_b200 as S
from tests._util import unpack_trains, flip_mask, spike_agreement

pytestmark = pytest.mark.gpu


def test_rpn_spike_rate_report_matches_oracle_format():
    W = O.reference_weights(seed=3)
    w = [W["shared_conv"], W["conv_cls"], W["conv_bbox"]]
    g = torch.Generator().manual_seed(9)
    feats = [torch.randn(2, 256, 10, 14, generator=g), torch.randn(2, 256, 5, 7, generator=g)]
    T = 8
    m = S.RPNHeadSNN(256, 3, T)
    with torch.no_grad():
        m.shared_conv.weight.copy_(w[0]); m.conv_cls.weight.copy_(w[1]); m.conv_bbox.weight.copy_(w[2])
    m = m.cuda(); m.record_spikes = True
    m([f.cuda() for f in feats])
    got = S.rpn_spike_rates_and_flops(m)
    want = O.rpn_head_rates(feats, *w, T, 3)
    assert len(got) == len(want) == 6
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert torch.equal(a[:, 1].cpu(), b[:, 1])                                   # FLOP constants (incl. the swapped *4)
        assert torch.allclose(a[:, 0].cpu(), b[:, 0], rtol=2e-3, atol=2e-6)          # rates (LI ones are signed means near 0)
    assert 0 < S.energy_ratio(got[0::3], T) < 1


def test_box_spike_rate_report_matches_oracle_format():
    torch.manual_seed(4)
    K, Hd, C, T, R = 1024, 256, 5, 12, 21
    w = [torch.randn(Hd, K) * 0.06, torch.randn(Hd, Hd) * 0.12, torch.randn(C, Hd) * 0.1, torch.randn(4 * C, Hd) * 0.1]
    x = torch.randn(R, K)
    m = S.FastRCNNPredictorSNNFull(K, Hd, C, T)
    with torch.no_grad():
        m.fc6.weight.copy_(w[0]); m.fc7.weight.copy_(w[1]); m.cls_score.weight.copy_(w[2]); m.bbox_pred.weight.copy_(w[3])
    m = m.cuda(); m.record_spikes = True
    m(x.cuda())
    got = S.box_spike_rates_and_flops(m)
    want = O.box_head_rates(x, *w, T)
    for a, b in zip(got, want):
        assert torch.equal(a[:, 1].cpu(), b[:, 1])
        bad = ((a[:, 0].cpu() - b[:, 0]).abs() > 2e-3 * b[:, 0].abs().max() + 1e-7).float().mean().item()
        assert bad <= 0.1        # rows hit by a near-threshold flip may move; the rest match


def test_attach_to_torchvision_faster_rcnn_runs_end_to_end():
    import torchvision
    from torchvision.models.detection import fasterrcnn_resnet50_fpn
    torch.manual_seed(0)
    model = fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, min_size=256, max_size=384)
    S.attach_snn_heads(model, num_steps_rpn=8, num_steps_detector=12, num_classes=9)
    assert isinstance(model.rpn.head, S.RPNHeadSNN) and isinstance(model.roi_heads.box_predictor, S.FastRCNNPredictorSNNFull)
    keys = set(model.state_dict().keys())
    for k in ("rpn.head.shared_conv.weight", "rpn.head.conv_cls.weight", "rpn.head.conv_bbox.weight",
              "roi_heads.box_predictor.fc6.weight", "roi_heads.box_predictor.fc7.weight",
              "roi_heads.box_predictor.cls_score.weight", "roi_heads.box_predictor.bbox_pred.weight"):
        assert k in keys
    model = model.cuda().eval()
    feats_seen = []
    model.rpn.head.register_forward_hook(lambda m, i, o: feats_seen.append(([f.detach().cpu() for f in i[0]], o)))
    imgs = [torch.rand(3, 240, 320, device="cuda"), torch.rand(3, 200, 300, device="cuda")]
    with torch.no_grad():
        det = model(imgs)
    assert len(det) == 2 and {"boxes", "labels", "scores"} <= set(det[0].keys())
    feats, (lo, bb) = feats_seen[0]
    assert len(feats) == 5
    h = model.rpn.head
    rlo, rbb = O.rpn_head_forward(feats, h.shared_conv.weight.detach().cpu(), h.conv_cls.weight.detach().cpu(),
                                  h.conv_bbox.weight.detach().cpu(), 8)
    for l in range(5):
        bad = ((lo[l].cpu() - rlo[l]).abs().amax(dim=1) > 1e-3 * max(rlo[l].abs().max().item(), 1e-6)).float().mean().item()
        assert bad <= 0.01, (l, bad)


@pytest.mark.parametrize("mode", ["fp32_exact", "fp16x2"])
@pytest.mark.parametrize("workload", ["cityscapes", "bdd"])
def test_full_size_level_shapes_parity(workload, mode):
    """BASELINE configs 1/2/3 at their real per-image sizes (one image): every FPN level of the
    Cityscapes (768x1536) and BDD (768x1376, ragged widths) shapes against the oracle."""
    levels = O.CITYSCAPES_LEVELS if workload == "cityscapes" else O.BDD_LEVELS
    C = 9 if workload == "cityscapes" else 5
    W = O.reference_weights(num_classes=C, seed=0)
    w = [W["shared_conv"], W["conv_cls"], W["conv_bbox"]]
    feats, rois = O.synthetic_inputs(levels, 1, rois_per_image=300)
    T = 8
    m = S.RPNHeadSNN(256, 3, T, mode=mode)
    with torch.no_grad():
        m.shared_conv.weight.copy_(w[0]); m.conv_cls.weight.copy_(w[1]); m.conv_bbox.weight.copy_(w[2])
    m = m.cuda(); m.record_spikes = True
    stats = {"workload": workload, "mode": mode, "rpn_levels": []}
    lo, bb = m([f.cuda() for f in feats])
    torch.cuda.synchronize()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    rlo, rbb, tr = O.rpn_head_forward(feats, *w, T, record=True)
    for l in range(len(levels)):
        trains = m.last_spike_trains[l].permute(0, 3, 1, 2)
        fm = flip_mask(trains, tr[l]["spk"], T)
        agree, unexplained = spike_agreement(trains, tr[l]["spk"], tr[l]["v_dec"], T)
        stats["rpn_levels"].append({"neurons": fm.numel(), "flipped_neurons": int(fm.sum()), "spike_agreement": agree,
                                    "flips_outside_1e-5_band": unexplained})
        assert unexplained == 0 and agree >= 0.999
        assert fm.float().mean().item() <= 1e-3, f"level {l}: flipped neurons {fm.float().mean().item()}"
        keep = ~fm.any(dim=1, keepdim=True)
        scale = rlo[l].abs().max().item()
        assert ((lo[l].cpu() - rlo[l]).abs() * keep).max().item() <= 1e-3 * scale
        assert ((bb[l].cpu() - rbb[l]).abs() * keep).max().item() <= 1e-3 * rbb[l].abs().max().item()
    b = S.FastRCNNPredictorSNNFull(12544, 1024, C, 12, mode=mode)
    with torch.no_grad():
        b.fc6.weight.copy_(W["fc6"]); b.fc7.weight.copy_(W["fc7"]); b.cls_score.weight.copy_(W["cls_score"])
        b.bbox_pred.weight.copy_(W["bbox_pred"])
    b = b.cuda(); b.record_spikes = True
    cls, dl = b(rois.cuda())
    rc, rd, trb = O.box_head_forward(rois, W["fc6"], W["fc7"], W["cls_score"], W["bbox_pred"], 12, record=True)
    f6 = flip_mask(b.last_spike_trains[0], trb["spk6"], 12); f7 = flip_mask(b.last_spike_trains[1], trb["spk7"], 12)
    assert f6.float().mean().item() <= 1e-3 and f7.float().mean().item() <= 1e-3
    stats["box"] = {"neurons_per_layer": f6.numel(), "lif6_flipped": int(f6.sum()), "lif7_flipped": int(f7.sum())}
    out_dir = os.environ.get("SNN_PARITY_STATS_DIR")          # profiles/: measured flip counts of both fp32-grade modes
    if out_dir:
        import json
        with open(os.path.join(out_dir, f"parity_stats_{workload}_{mode}.json"), "w") as f:
            json.dump(stats, f)
    keep = ~(f6.any(dim=1) | f7.any(dim=1)).unsqueeze(1)
    assert ((cls.cpu() - rc).abs() * keep).max().item() <= 1e-3 * rc.abs().max().item()
    assert ((dl.cpu() - rd).abs() * keep).max().item() <= 1e-3 * rd.abs().max().item()


def test_fast_postprocessing_keeps_the_detections_of_the_stock_path():
    """attach_fast_postprocessing: proposals selected from the head's NCHW outputs + decode of the
    selected anchors only must give the detections of torchvision's own RegionProposalNetwork.forward."""
    from torchvision.models.detection import fasterrcnn_resnet50_fpn
    torch.manual_seed(1)
    model = fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, min_size=256, max_size=384,
                                    box_score_thresh=0.0)
    S.attach_snn_heads(model, num_steps_rpn=8, num_steps_detector=12, num_classes=9)
    model = model.cuda().eval()
    imgs = [torch.rand(3, 240, 320, device="cuda"), torch.rand(3, 200, 300, device="cuda")]
    props_seen = []
    model.roi_heads.register_forward_pre_hook(lambda m, args: props_seen.append([p.detach().cpu() for p in args[1]]))
    with torch.no_grad():
        want = model(imgs)
        S.attach_fast_postprocessing(model)
        got = model(imgs)
    ref_props, fast_props = props_seen
    for a, b in zip(ref_props, fast_props):               # the proposals handed to the RoI heads
        assert a.shape == b.shape and a.shape[0] > 0
        assert torch.allclose(a, b, atol=2e-3, rtol=1e-5)
    for a, b in zip(want, got):
        assert a["boxes"].shape == b["boxes"].shape
        assert torch.allclose(a["boxes"], b["boxes"], atol=5e-2)
        assert torch.equal(a["labels"], b["labels"])
