"""Host-side behaviour of the drop-in modules that needs no GPU: the training-mode guard, the prepared-weight cache
hooks and the error paths that must fire before any device work."""
import pytest
import torch

import snn_automotive_object_detection_b200 as S


def test_training_mode_forward_is_refused_loudly():
    """The reference trains through these heads (train.py:149-200, Norse surrogate gradients); the CUDA heads build no
    autograd graph, so a training-mode forward with gradients enabled must raise instead of returning detached outputs."""
    rpn = S.RPNHeadSNN(256, 3, 8)
    box = S.FastRCNNPredictorSNNFull(256, 128, 3, 12)
    assert rpn.training and box.training
    with pytest.raises(RuntimeError, match="inference-only"):
        rpn([torch.randn(1, 256, 4, 4)])
    with pytest.raises(RuntimeError, match="inference-only"):
        box(torch.randn(2, 256))
    # under no_grad (or in eval mode) the guard lets the call through to the device check
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        rpn([torch.randn(1, 256, 4, 4)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        box.eval()(torch.randn(2, 256))
    # frozen heads fed by a trainable backbone would silently cut the graph as well
    for p in rpn.parameters():
        p.requires_grad_(False)
    with pytest.raises(RuntimeError, match="inference-only"):
        rpn([torch.randn(1, 256, 4, 4, requires_grad=True)])


def test_prepared_weight_cache_is_invalidated_by_load_and_apply():
    box = S.FastRCNNPredictorSNNFull(256, 128, 3, 12)
    assert len(box._prepared()) == 2
    for hook in (lambda: box.load_state_dict(box.state_dict()), lambda: box.float(), lambda: box.to("cpu"),
                 box.invalidate_weight_cache):
        for w in box._prepared():
            w.key, w.buf = ("stale",), object()
        hook()
        assert all(w.key is None and w.buf is None for w in box._prepared())


def test_state_dict_keys_are_the_references():
    rpn = S.RPNHeadSNN(256, 3, 8)
    box = S.FastRCNNPredictorSNNFull(12544, 1024, 9, 12)
    assert sorted(rpn.state_dict()) == ["conv_bbox.weight", "conv_cls.weight", "shared_conv.weight"]
    assert sorted(box.state_dict()) == ["bbox_pred.weight", "cls_score.weight", "fc6.weight", "fc7.weight"]
    assert float(rpn.p_enc.v_th) == 0.25 and rpn.num_steps == 8 and box.num_steps == 12
