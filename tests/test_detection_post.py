"""The "next" rows against goldens produced by the UNMODIFIED reference classes
(oracle/gen_golden_post.py): vectorised RoIHeadsSNN.postprocess_detections (CPU and GPU) and the proposal
selection fed from the RPN head's native NCHW layout (GPU, through the C ABI)."""
import os

import numpy as np
import pytest
import torch
from torchvision.models.detection import _utils as det_utils

from snn_automotive_object_detection_b200 import detection_post as DP


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _post_inputs(g, device="cpu"):
    per_img = [int(v) for v in g["per_img"]]
    props = [torch.from_numpy(g[f"props{i}"]).to(device) for i in range(len(per_img))]
    shapes = [tuple(int(v) for v in s) for s in g["shapes"]]
    return torch.from_numpy(g["logits"]).to(device), torch.from_numpy(g["reg"]).to(device), props, shapes


def _check_post(g, out, exact):
    boxes, scores, labels, all_scores, all_boxes = out
    for i in range(len(boxes)):
        assert torch.equal(labels[i].cpu(), torch.from_numpy(g[f"labels{i}"])), "same detections in the same order"
        if exact:
            assert torch.equal(boxes[i].cpu(), torch.from_numpy(g[f"boxes{i}"]))
            assert torch.equal(scores[i].cpu(), torch.from_numpy(g[f"scores{i}"]))
            assert torch.equal(all_scores[i].cpu(), torch.from_numpy(g[f"all_scores{i}"]))
            assert torch.equal(all_boxes[i].cpu(), torch.from_numpy(g[f"all_boxes{i}"]))
        else:
            assert torch.allclose(boxes[i].cpu(), torch.from_numpy(g[f"boxes{i}"]), atol=2e-3)
            assert torch.allclose(scores[i].cpu(), torch.from_numpy(g[f"scores{i}"]), atol=1e-6)
            assert torch.allclose(all_boxes[i].cpu(), torch.from_numpy(g[f"all_boxes{i}"]), atol=2e-3)


def test_vectorised_postprocess_equals_the_reference_loop(golden_dir):
    g = _load(golden_dir, "post_detections")
    logits, reg, props, shapes = _post_inputs(g)
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))                    # roi_heads.py:938-940 default
    out = DP.postprocess_detections(logits, reg, props, shapes, coder, float(g["score_thresh"]),
                                    float(g["nms_thresh"]), int(g["detections_per_img"]))
    n_bg = [int((out[2][i] == 0).sum()) for i in range(2)]
    assert min(n_bg) > 0 and all(int((out[2][i] > 0).sum()) > 0 for i in range(2)), "both objects and background kept"
    _check_post(g, out, exact=True)


def test_patch_binds_the_reference_signature(golden_dir):
    g = _load(golden_dir, "post_detections")
    logits, reg, props, shapes = _post_inputs(g)

    class Holder:
        box_coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
        score_thresh, nms_thresh, detections_per_img = 0.4, 0.5, 100

    h = DP.patch_postprocess(Holder())
    _check_post(g, h.postprocess_detections(logits, reg, props, shapes), exact=True)


def test_nchw_index_maps_to_the_reference_flattening():
    A, H, W = 3, 5, 7
    x = torch.arange(A * H * W).view(1, A, H, W)
    ref = x.view(1, -1, 1, H, W).permute(0, 3, 4, 1, 2).reshape(-1)        # permute_and_flatten, rpn.py:256-258
    idx = torch.arange(A * H * W)
    assert torch.equal(ref[DP.nchw_index_to_reference_order(idx, A, H, W)], idx)


@pytest.mark.gpu
def test_vectorised_postprocess_on_gpu(golden_dir):
    g = _load(golden_dir, "post_detections")
    logits, reg, props, shapes = _post_inputs(g, "cuda")
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    out = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.4, 0.5, 100)
    _check_post(g, out, exact=False)


@pytest.mark.gpu
def test_rpn_proposals_from_native_layout_match_the_reference(golden_dir):
    g = _load(golden_dir, "post_rpn")
    L = len(g["levels"])
    obj = [torch.from_numpy(g[f"obj{l}"]).cuda() for l in range(L)]
    dlt = [torch.from_numpy(g[f"dlt{l}"]).cuda() for l in range(L)]
    cells = [torch.from_numpy(g[f"cell{l}"]) for l in range(L)]
    img = tuple(int(v) for v in g["img"])
    strides = [(img[0] // int(h), img[1] // int(w)) for (h, w) in g["levels"]]
    props, probs, levels, ref_index = DP.rpn_select_proposals(obj, dlt, cells, strides, int(g["pre_nms_top_n"]))
    torch.cuda.synchronize()
    N = props.shape[0]
    for i in range(N):
        # pre-NMS: same selected set, same decoded boxes / probabilities (top-k order may differ only between ties)
        want_p, want_o = torch.from_numpy(g[f"pre_props{i}"]), torch.from_numpy(g[f"pre_obj{i}"])
        assert props[i].shape == want_p.shape
        assert torch.allclose(probs[i].cpu(), want_o, atol=1e-6)           # both sorted by objectness per level
        assert torch.allclose(props[i].cpu(), want_p, atol=2e-3, rtol=1e-5)
        assert ref_index[i].unique().numel() == ref_index[i].numel()
    image_sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    boxes, scores, pre = DP.filter_selected(props, probs, levels, image_sizes, float(g["min_size"]), 0.0,
                                            float(g["nms_thresh"]), int(g["post_nms_top_n"]))
    for i in range(N):
        want = torch.from_numpy(g[f"boxes{i}"])
        assert boxes[i].shape == want.shape
        assert torch.allclose(boxes[i].cpu(), want, atol=2e-3, rtol=1e-5)
        assert torch.equal(pre[i]["proposals"], props[i])


@pytest.mark.gpu
def test_ref_index_points_at_the_reference_anchor(golden_dir):
    # the anchor index reported for a selected entry is its position in the reference's (level, h, w, a) order
    g = _load(golden_dir, "post_rpn")
    L = len(g["levels"])
    obj = [torch.from_numpy(g[f"obj{l}"]).cuda() for l in range(L)]
    dlt = [torch.from_numpy(g[f"dlt{l}"]).cuda() for l in range(L)]
    cells = [torch.from_numpy(g[f"cell{l}"]) for l in range(L)]
    img = tuple(int(v) for v in g["img"])
    strides = [(img[0] // int(h), img[1] // int(w)) for (h, w) in g["levels"]]
    props, probs, levels, ref_index = DP.rpn_select_proposals(obj, dlt, cells, strides, 50)
    flat = torch.cat([o.permute(0, 2, 3, 1).reshape(o.shape[0], -1) for o in obj], dim=1)   # (level, h, w, a)
    got = torch.sigmoid(torch.gather(flat, 1, ref_index))
    assert torch.allclose(got, probs, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["unroll2", "unroll4"])
@pytest.mark.parametrize("names,sampling", [(["0", "1", "2", "3"], 2), (["0"], 2), (["0", "1", "2", "3"], 1), (["1"], 0)])
def test_fused_roi_align_encoder_matches_torchvision_pool_then_encode(names, sampling, kernel):
    """RoIAlign fused with the box head's encoder against torchvision's MultiScaleRoIAlign followed by
    the head's own encoder, and the box head outputs on both inputs; multi-level and single-level poolers, fixed 2 x 2 and
    1 x 1 sampling grids and torchvision's adaptive grid (sampling_ratio 0), both unroll variants of the kernel."""
    from collections import OrderedDict
    from torchvision.ops import MultiScaleRoIAlign
    from oracle import snn_oracle as O
    import snn_automotive_object_detection_b200 as S
    from snn_automotive_object_detection_b200 import _lib
    from snn_automotive_object_detection_b200.heads import unpack_trains
    torch.manual_seed(3)
    N, C, T = 2, 256, 12
    img = (256, 384)
    feats = OrderedDict((str(i), torch.randn(N, C, img[0] // s, img[1] // s, device="cuda")) for i, s in enumerate((4, 8, 16, 32)))
    feats["pool"] = torch.randn(N, C, 4, 6, device="cuda")
    props = []
    for _ in range(N):
        xy = torch.rand(60, 2, device="cuda") * torch.tensor([300.0, 200.0], device="cuda")
        wh = torch.rand(60, 2, device="cuda") ** 2 * torch.tensor([300.0, 220.0], device="cuda") + 2.0
        props.append(torch.cat([xy, xy + wh], dim=1))
    props[0][0] = torch.tensor([-20.0, -10.0, 500.0, 300.0], device="cuda")      # sticks out of the image
    props[1][1] = torch.tensor([370.0, 250.0, 420.0, 300.0], device="cuda")      # almost entirely outside
    props[1][2] = torch.tensor([100.0, 100.0, 100.5, 100.5], device="cuda")      # smaller than one feature pixel
    shapes = [img, (240, 360)]
    pooler = MultiScaleRoIAlign(featmap_names=names, output_size=7, sampling_ratio=sampling)
    want = pooler(feats, props, shapes)                                           # [R, C, 7, 7]
    fused = S.FusedRoIAlignEncoder.from_pooler(pooler, T)
    fused.return_pooled = True
    lib = _lib.load()
    lib.snn_set_roi_kernel(1 if kernel == "unroll4" else 0)
    try:
        enc = fused(feats, props, shapes)
        torch.cuda.synchronize()
    finally:
        lib.snn_set_roi_kernel(0)
    got = fused.last_pooled.view_as(want)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5), (got - want).abs().max().item()
    # the words are the encoder's spike trains of the pooled values (bit-exact w.r.t. the kernel's own pooled values)
    ref_spk = torch.stack(O.encoder_spikes(fused.last_pooled.cpu(), T - 1))       # [T-1, R, K]
    assert torch.equal(unpack_trains(enc.words.cpu(), T - 1).float(), ref_spk)
    # the head: identical outputs from the words and from the kernel's own pooled values (same spike trains in, bit
    # for bit); against torchvision's pooled tensor every RoI whose words agree gives identical outputs, and the RoIs
    # with a differing word (an input within RoIAlign's ~1e-7 rounding of an encoder threshold) stay a small share
    head = S.FastRCNNPredictorSNNFull(C * 49, 1024, 9, T).cuda().eval()
    cls_b, box_b = head(enc)
    cls_c, box_c = head(fused.last_pooled.view_as(want))
    assert torch.equal(cls_b, cls_c) and torch.equal(box_b, box_c)
    cls_a, box_a = head(want)
    tv_words = torch.stack(O.encoder_spikes(want.flatten(1).cpu(), T - 1))
    same = (tv_words == ref_spk).all(dim=0).all(dim=1)
    assert same.float().mean().item() >= 0.8, same.float().mean().item()
    assert torch.equal(cls_a.cpu()[same], cls_b.cpu()[same]) and torch.equal(box_a.cpu()[same], box_b.cpu()[same])


@pytest.mark.gpu
@pytest.mark.parametrize("k", [400, 1000, 3000])
def test_topk_ties_are_broken_in_the_references_order(k):
    """Pixels where no shared_lif neuron spiked have exactly-zero membranes for every anchor, so large tie groups at
    the k boundary are real.  The reference takes top-k on the (H, W, A)-flattened logits (rpn.py:248-259, 468-472);
    the selection here must be "largest first, ties by the lowest reference index" -- a stable descending sort of the
    reference-ordered tensor.  k <= 2048 runs the library's radix-select + sort kernels over all (level, image)
    segments at once (k = 400 cuts the zero plateau of the large level; the small level has fewer anchors than k and is
    sorted whole); k = 3000 takes the key tensor + torch.topk path."""
    import snn_automotive_object_detection_b200 as S
    torch.manual_seed(11)
    N, A = 3, 3
    shapes = [(40, 50), (5, 7)]
    logits, deltas = [], []
    for (H, W) in shapes:
        lg = torch.zeros(N, A, H, W)
        hot = torch.rand(N, A, H, W) < 0.02                  # 2 % distinct positive values, the rest exact ties at 0
        lg[hot] = torch.rand(int(hot.sum())) + 0.1
        neg = (torch.rand(N, A, H, W) < 0.3) & ~hot
        lg[neg] = -torch.rand(int(neg.sum())) - 0.1
        lg[0, 1, 3, 4] = -0.0
        lg[1, 0, 0, 0] = float("inf")
        lg[2, 2, 1, 1] = -float("inf")
        lg[1, 2, 2:4, :] = 0.25                              # a tie group of positive values as well
        logits.append(lg); deltas.append(0.1 * torch.randn(N, 4 * A, H, W))
    cell = torch.tensor([[-16., -8., 16., 8.], [-11., -11., 11., 11.], [-8., -16., 8., 16.]])
    _, probs, levels, ref_index = S.rpn_select_proposals([t.cuda() for t in logits], [t.cuda() for t in deltas],
                                                         [cell, cell], [(16, 16), (32, 32)], k)
    want, off = [], 0
    for lg in logits:
        flat = lg.permute(0, 2, 3, 1).reshape(N, -1)         # the reference's order
        kk = min(k, flat.shape[1])
        want.append(torch.sort(flat, dim=1, descending=True, stable=True)[1][:, :kk] + off)
        off += flat.shape[1]
    want = torch.cat(want, dim=1)
    assert ref_index.shape == want.shape and torch.equal(ref_index.cpu(), want)
    allf = torch.cat([lg.permute(0, 2, 3, 1).reshape(N, -1) for lg in logits], dim=1)
    assert torch.allclose(probs.cpu(), torch.sigmoid(torch.gather(allf, 1, want)), atol=1e-6)
    assert levels.shape == want.shape and int(levels[0, -1]) == 1


def _random_detector_outputs(seed, rois, C, logit_scale, device="cuda", duplicate_rows=0, image_hw=(768, 1536)):
    """Class logits, box regression and proposals with the statistics of a trained detector's outputs: confident scores
    (logit_scale), proposals spread over the image with heavy overlap, small regression deltas."""
    g = torch.Generator().manual_seed(seed)
    ih, iw = image_hw
    logits, reg, props = [], [], []
    for R in rois:
        lg = torch.randn(R, C, generator=g) * logit_scale
        ctr = torch.rand(R, 2, generator=g) * torch.tensor([iw * 1.0, ih * 1.0])
        ctr = (ctr / 96).round() * 96 + torch.randn(R, 2, generator=g) * 12      # clusters of overlapping proposals
        wh = (torch.rand(R, 2, generator=g) * 0.8 + 0.6) * torch.tensor([120.0, 90.0])
        pr = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
        rg = torch.randn(R, 4 * C, generator=g) * torch.tensor([1.0, 1.0, 0.5, 0.5]).repeat(C)
        if duplicate_rows and R > 2 * duplicate_rows:           # exact ties: identical scores and boxes
            lg[duplicate_rows:2 * duplicate_rows] = lg[:duplicate_rows]
            pr[duplicate_rows:2 * duplicate_rows] = pr[:duplicate_rows]
            rg[duplicate_rows:2 * duplicate_rows] = rg[:duplicate_rows]
        logits.append(lg); reg.append(rg); props.append(pr.to(device))
    return torch.cat(logits).to(device), torch.cat(reg).to(device), props, [image_hw] * len(rois)


@pytest.mark.gpu
@pytest.mark.parametrize("name,rois,C,scale,thresh,per_img,dups", [
    ("bench_shape", [1000, 1000], 9, 3.0, 0.4, 100, 0),
    ("low_threshold_many_candidates", [1000, 1000], 9, 1.0, 0.05, 100, 0),     # > 5000 boxes: per-class NMS, no shift
    ("bdd_5_classes", [1000], 5, 3.0, 0.4, 100, 0),
    ("ragged_and_empty", [0, 7, 513], 5, 3.0, 0.3, 100, 0),
    ("two_classes", [300, 20], 2, 2.0, 0.5, 100, 0),
    ("exact_ties", [400, 400], 9, 3.0, 0.4, 100, 37),
    ("nothing_passes", [200], 9, 0.1, 0.9, 100, 0),
    ("few_detections_kept", [1000], 9, 3.0, 0.2, 5, 0),
    ("max_candidates", [1024], 9, 2.0, 0.01, 300, 0),                          # 8192 (RoI, class) pairs: the kernel's limit
])
def test_postprocess_kernel_equals_the_torchvision_path(name, rois, C, scale, thresh, per_img, dups):
    """SURVEY 8f-4 on the device: `snn_det_postprocess` (one launch per batch) against the torch/torchvision ops of the
    reference's own function on the same softmax scores and decoded boxes -- the same detections, in the same order,
    bit for bit: boxes, scores, labels, all_scores, all_boxes."""
    logits, reg, props, shapes = _random_detector_outputs(sum(map(ord, name)), rois, C, scale, duplicate_rows=dups)
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    want = DP.postprocess_detections(logits, reg, props, shapes, coder, thresh, 0.5, per_img, use_kernel=False)
    got = DP.postprocess_detections(logits, reg, props, shapes, coder, thresh, 0.5, per_img, use_kernel=True)
    auto = DP.postprocess_detections(logits, reg, props, shapes, coder, thresh, 0.5, per_img)
    n_obj = sum(int((l > 0).sum()) for l in want[2])
    n_bg = sum(int((l == 0).sum()) for l in want[2])
    if name == "nothing_passes":
        assert n_obj == 0 and n_bg > 0
    elif name in ("low_threshold_many_candidates", "max_candidates"):
        assert n_obj > 0 and n_bg == 0, (n_obj, n_bg)         # every RoI has a class over the threshold
    else:
        assert n_obj > 0 and n_bg > 0, (n_obj, n_bg)
    for out in (got, auto):
        for k in range(5):
            assert len(out[k]) == len(want[k]) == len(rois)
            for i in range(len(rois)):
                assert out[k][i].shape == want[k][i].shape, (name, k, i, out[k][i].shape, want[k][i].shape)
                assert out[k][i].dtype == want[k][i].dtype
                assert torch.equal(out[k][i], want[k][i]), (name, k, i)


@pytest.mark.gpu
def test_postprocess_kernel_on_the_reference_golden(golden_dir):
    g = _load(golden_dir, "post_detections")
    logits, reg, props, shapes = _post_inputs(g, "cuda")
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    out = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.4, 0.5, 100, use_kernel=True)
    _check_post(g, out, exact=False)


@pytest.mark.gpu
def test_postprocess_kernel_limits_fall_back_or_fail_loudly():
    logits, reg, props, shapes = _random_detector_outputs(5, [1200], 9, 2.0)     # 9600 pairs > 8192
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    want = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.4, 0.5, 100, use_kernel=False)
    auto = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.4, 0.5, 100)       # torch ops: over the limit
    assert all(torch.equal(a, b) for a, b in zip(auto[0], want[0]))
    with pytest.raises(RuntimeError, match="candidates"):
        DP.postprocess_detections(logits, reg, props, shapes, coder, 0.4, 0.5, 100, use_kernel=True)


def _random_selected_proposals(seed, N, level_sizes, image_hw, saturated=0, device="cuda"):
    """Level-major selected proposals as rpn_select_proposals returns them: per level, objectness descending; boxes
    clustered so that NMS at 0.7 has work to do; some boxes outside the image or degenerate."""
    g = torch.Generator().manual_seed(seed)
    ih, iw = image_hw
    props, probs, levels = [], [], []
    for l, k in enumerate(level_sizes):
        size = 32.0 * 2 ** l
        ctr = torch.rand(N, k, 2, generator=g) * torch.tensor([iw * 1.1, ih * 1.1]) - torch.tensor([iw * 0.05, ih * 0.05])
        ctr = (ctr / (size / 2)).round() * (size / 2) + torch.randn(N, k, 2, generator=g) * size * 0.08
        wh = (torch.rand(N, k, 2, generator=g) * 0.6 + 0.7) * size
        wh[:, : max(1, k // 50)] *= 1e-5                                   # a few boxes below min_size
        props.append(torch.cat([ctr - wh / 2, ctr + wh / 2], dim=2))
        pr = torch.sigmoid(torch.randn(N, k, generator=g) * 3).sort(dim=1, descending=True)[0]
        if saturated:
            pr[:, :saturated] = 1.0                                       # exact ties at the top (a saturated sigmoid)
        probs.append(pr)
        levels.append(torch.full((N, k), l, dtype=torch.int64))
    return torch.cat(props, 1).to(device), torch.cat(probs, 1).to(device), torch.cat(levels, 1).to(device)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,level_sizes,post_n,score_thresh,saturated", [
    ("bench_shape", 2, [1000, 1000, 1000, 1000, 864], 1000, 0.0, 0),
    ("training_sizes_over_5000_boxes", 2, [2000, 2000, 2000, 1152, 288], 2000, 0.0, 0),     # per-class NMS, no shift
    ("score_threshold", 3, [1000, 1000, 500, 100, 7], 300, 0.3, 0),
    ("saturated_scores_tie", 2, [600, 600, 600], 1000, 0.0, 25),
    ("one_level", 1, [1500], 100, 0.0, 0),
    ("nothing_left", 2, [50, 50], 100, 1.5, 0),
])
def test_rpn_filter_kernel_equals_the_torchvision_path(name, N, level_sizes, post_n, score_thresh, saturated):
    """SURVEY 8f-1, the tail of filter_proposals (rpn.py:493-525): `snn_rpn_nms` (one launch, one block per (image,
    level), the last block of an image merges) against clip / remove_small_boxes / threshold / torchvision batched_nms /
    top-n on the same inputs -- the same proposals in the same order, bit for bit."""
    shapes = [(768, 1536), (700, 1400), (768, 1365)][:N]
    props, probs, levels = _random_selected_proposals(sum(map(ord, name)), N, level_sizes, (768, 1536), saturated)
    want = DP.filter_selected(props, probs, levels, shapes, 1e-3, score_thresh, 0.7, post_n, use_kernel=False)
    got = DP.filter_selected(props, probs, levels, shapes, 1e-3, score_thresh, 0.7, post_n, level_sizes=level_sizes,
                             use_kernel=True)
    auto = DP.filter_selected(props, probs, levels, shapes, 1e-3, score_thresh, 0.7, post_n, level_sizes=level_sizes)
    n_kept = [int(b.shape[0]) for b in want[0]]
    if name == "nothing_left":
        assert n_kept == [0] * N
    else:
        assert min(n_kept) > 0 and (name != "bench_shape" or min(n_kept) > 100), n_kept
    for out in (got, auto):
        for i in range(N):
            assert out[0][i].shape == want[0][i].shape, (name, i, out[0][i].shape, want[0][i].shape)
            assert torch.equal(out[1][i], want[1][i]), (name, i, "scores")
            assert torch.equal(out[0][i], want[0][i]), (name, i, "boxes")


@pytest.mark.gpu
@pytest.mark.parametrize("thresh,per_img", [(0.05, 100), (0.4, 100), (0.2, 7)])
def test_torchvision_roi_heads_postprocess_patched_equals_the_stock_method(thresh, per_img):
    """A stock torchvision RoIHeads (what attach_snn_heads leaves on a torchvision Faster R-CNN): its own
    postprocess_detections against the patched one (snn_det_postprocess, object rows) -- same detections, same order."""
    from torchvision.models.detection import fasterrcnn_resnet50_fpn
    model = fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, box_score_thresh=thresh,
                                    box_detections_per_img=per_img)
    rh = model.roi_heads
    logits, reg, props, shapes = _random_detector_outputs(77, [1000, 640], 9, 2.0)
    want = rh.postprocess_detections(logits, reg, props, shapes)
    DP.patch_postprocess_torchvision(rh)
    got = rh.postprocess_detections(logits, reg, props, shapes)
    assert sum(int(b.shape[0]) for b in want[0]) > 0
    for k in range(3):
        for a, b in zip(want[k], got[k]):
            assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), (k, a.shape, b.shape)
    cpu = rh.postprocess_detections(logits.cpu(), reg.cpu(), [p.cpu() for p in props], shapes)     # torchvision's method
    assert all(a.shape == b.shape for a, b in zip(cpu[2], want[2]))


@pytest.mark.gpu
def test_post_kernels_with_more_images_than_one_launch_holds():
    """Both kernels keep the images' geometry in their parameter block, 64 images a launch: 70 images take two."""
    N = 70
    rois = [12 + (i * 7) % 30 for i in range(N)]
    logits, reg, props, shapes = _random_detector_outputs(3, rois, 5, 2.0)
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    want = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.3, 0.5, 10, use_kernel=False)
    got = DP.postprocess_detections(logits, reg, props, shapes, coder, 0.3, 0.5, 10, use_kernel=True)
    for k in range(5):
        assert all(torch.equal(a, b) for a, b in zip(want[k], got[k])), k
    level_sizes = [40, 0, 25]                                               # and a level without entries
    p, s, lv = _random_selected_proposals(9, N, level_sizes, (256, 320))
    shapes = [(256, 320)] * N
    want = DP.filter_selected(p, s, lv, shapes, 1e-3, 0.0, 0.7, 30, use_kernel=False)
    got = DP.filter_selected(p, s, lv, shapes, 1e-3, 0.0, 0.7, 30, level_sizes=level_sizes, use_kernel=True)
    for k in range(2):
        assert all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(want[k], got[k])), k


def _threshold_pairs(seed, n_pairs, ratio, image_hw, device="cuda"):
    """Pairs of boxes whose IoU is `ratio` in exact arithmetic (one box is the other cut to `ratio` of its height), at
    random scales and translations: in fp32 every pair lands within a few ulps of the NMS threshold, on either side,
    so the kept set depends on every rounding of the IoU (the FMUL / FFMA / division sequence torchvision compiles to)."""
    g = torch.Generator().manual_seed(seed)
    ih, iw = image_hw
    s = torch.rand(n_pairs, 1, generator=g) * 40 + 3
    t = torch.rand(n_pairs, 2, generator=g) * torch.tensor([iw - 60.0, ih - 60.0])
    a = torch.cat([t, t + s], dim=1)
    b = torch.cat([t, t[:, :1] + s, t[:, 1:] + s * ratio], dim=1)
    return torch.stack([a, b], dim=1).reshape(-1, 4).to(device)             # a0, b0, a1, b1, ...


@pytest.mark.gpu
def test_nms_decisions_on_the_threshold_follow_torchvisions_rounding():
    # detector post-processing, NMS 0.5: 600 pairs with IoU 0.5 +- ulps, one object class, distinct scores
    boxes = _threshold_pairs(1, 600, 0.5, (768, 1536))
    R = boxes.shape[0]
    g = torch.Generator().manual_seed(2)
    logits = torch.stack([torch.zeros(R), torch.rand(R, generator=g) * 4 + 1], dim=1).cuda()
    reg = torch.zeros(R, 8, device="cuda")
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    want = DP.postprocess_detections(logits, reg, [boxes], [(768, 1536)], coder, 0.4, 0.5, 2000, use_kernel=False)
    got = DP.postprocess_detections(logits, reg, [boxes], [(768, 1536)], coder, 0.4, 0.5, 2000, use_kernel=True)
    kept = int(want[0][0].shape[0])
    assert R * 0.55 < kept < R * 0.95, kept                 # a good share of the pairs falls on each side
    for k in range(5):
        assert torch.equal(want[k][0], got[k][0]), k
    # proposal filter, NMS 0.7, two levels (the second one shifted by the coordinate trick, which rounds the boxes)
    props = _threshold_pairs(3, 500, 0.7, (768, 1536)).reshape(1, -1, 4)
    K = props.shape[1]
    probs = torch.rand(1, K, generator=g).sort(dim=1, descending=True)[0].cuda()
    sizes = [K // 2, K - K // 2]
    probs = torch.cat([probs[:, 0::2], probs[:, 1::2]], dim=1).contiguous()         # descending inside each level
    levels = torch.cat([torch.zeros(1, sizes[0]), torch.ones(1, sizes[1])], dim=1).long().cuda()
    want = DP.filter_selected(props, probs, levels, [(768, 1536)], 1e-3, 0.0, 0.7, 2000, use_kernel=False)
    got = DP.filter_selected(props, probs, levels, [(768, 1536)], 1e-3, 0.0, 0.7, 2000, level_sizes=sizes, use_kernel=True)
    kept = int(want[0][0].shape[0])
    assert K * 0.55 < kept < K * 0.95, kept
    assert torch.equal(want[0][0], got[0][0]) and torch.equal(want[1][0], got[1][0])
