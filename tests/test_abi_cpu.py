"""The C-ABI library loads without a GPU and exports every symbol include/snn_heads.h declares; the host-only
entry points (no kernel launch) behave as documented.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from snn_automotive_object_detection_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "snn_heads.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snn_[a-z0-9_]+)\s*\(", text)) - {"snn_stream_t"})


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from snn_automotive_object_detection_b200.csrc import build
        build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/snn_heads.h but not exported by the library"
    assert sorted(_lib.EXPORTS) == names, "the ctypes binding list and the header disagree"


def test_abi_version_matches_the_header(lib):
    want = int(re.search(r"#define SNN_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.snn_version() == want == _lib.EXPECTED_ABI


def test_a_stale_library_is_refused(lib, monkeypatch):
    """load() compares snn_version() with the ABI its argtypes describe before binding anything."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "EXPECTED_ABI", _lib.EXPECTED_ABI + 1)
    with pytest.raises(RuntimeError, match="stale"):
        _lib.load()


def test_host_only_helpers(lib):
    assert [lib.snn_train_word_bytes(t) for t in (1, 8, 9, 16, 17, 32)] == [1, 1, 2, 2, 4, 4]
    modes = {0: 3, 1: 1, 2: 2, 3: 2, 4: 1}
    for m, pieces in modes.items():
        assert lib.snn_mode_pieces(m) == pieces
        assert lib.snn_prepared_weight_bytes(1024, 12544, m) == pieces * 1024 * 12544 * 2 + 1024 * 4
    assert lib.snn_mode_pieces(99) == 0
    assert set(_lib.MODES.values()) == set(modes)


def test_bad_arguments_are_reported_without_touching_a_device(lib):
    rc = lib.snn_prepare_fc_weights(None, 4, 64, 0, None, None)
    assert rc == -1 and b"bad argument" in lib.snn_last_error()
    rc = lib.snn_rpn_decode_selected(None, None, None, None, None, None, None, None, 1, 1, 3, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.snn_last_error()
    with pytest.raises(ValueError):
        _lib.mode_id("int8")


def test_encoder_comparator_table_against_the_oracle_encoder(lib):
    """The encoders evaluate lif_current_encoder as a comparator bank (include/snn_heads.h: snn_encoder_table).
    Its thresholds are pinned here against the oracle's step-by-step encoder: thresholds[n] is the SMALLEST fp32
    input whose first spike comes at step <= n, the first-spike step is monotone in the input over every fp32
    value between the smallest and the largest threshold, and the word the table yields equals the simulated
    train for all those inputs."""
    import numpy as np
    import torch
    from oracle import snn_oracle as O
    thr = (ctypes.c_float * 33)()
    dl = (ctypes.c_uint * 33)()
    lib.snn_encoder_table(thr, dl)
    thr = np.array(thr[:], dtype=np.float32)
    dl = np.array(dl[:], dtype=np.uint32)
    T = 32

    def first_spike(x):
        z = torch.stack(O.encoder_spikes(torch.from_numpy(x), T)).numpy() > 0       # [T, n]
        return np.where(z.any(0), z.argmax(0) + 1, T + 1), z

    for n in range(1, 33):
        x = np.array([np.nextafter(thr[n], np.float32(0)), thr[n]], dtype=np.float32)
        f, _ = first_spike(x)
        assert f[0] > n and f[1] <= n, f"threshold {n}: {thr[n]!r} is not the smallest input spiking by step {n}"
        full = sum(1 << t for t in range(n - 1, 32, n))
        nxt = sum(1 << t for t in range(n, 32, n + 1)) if n < 32 else 0
        assert int(dl[n]) == full ^ nxt
    assert np.all(np.diff(thr[1:]) < 0)
    # every fp32 value in [just below thr[32], just above thr[1]]: monotone first-spike step, table word == simulation
    lo = int(np.float32(thr[32]).view(np.uint32)) - 4096
    hi = int(np.float32(thr[1]).view(np.uint32)) + 4096
    prev_last = T + 1
    for s0 in range(lo, hi + 1, 1 << 22):
        x = np.arange(s0, min(s0 + (1 << 22), hi + 1), dtype=np.uint32).view(np.float32)
        f, z = first_spike(x)
        assert f[0] <= prev_last and np.all(np.diff(f) <= 0)
        prev_last = f[-1]
        sim = np.zeros(x.shape, dtype=np.uint32)
        for t in range(T):
            sim |= z[t].astype(np.uint32) << np.uint32(t)
        tab = np.zeros(x.shape, dtype=np.uint32)
        for n in range(1, 33):
            tab ^= np.where(x >= thr[n], dl[n], np.uint32(0))
        assert np.array_equal(sim, tab)


def test_workspace_sizes_are_host_only_and_cover_the_carried_state(lib):
    """The workspace queries run without a device.  Up to 17 steps a head needs its encoder words and spike trains;
    beyond that the time axis runs in passes and the workspace also holds the carried neuron state (16 B per neuron)."""
    H = (ctypes.c_int * 2)(24, 12)
    W = (ctypes.c_int * 2)(48, 24)
    neurons = 2 * (24 * 48 + 12 * 24) * 256
    w8 = lib.snn_rpn_head_workspace_bytes(H, W, 2, 2, 256, 8, 3)
    w17 = lib.snn_rpn_head_workspace_bytes(H, W, 2, 2, 256, 17, 3)
    w18 = lib.snn_rpn_head_workspace_bytes(H, W, 2, 2, 256, 18, 3)
    assert neurons * 2 <= w8 <= neurons * 2 + 16384                 # 1-byte encoder words + 1-byte trains
    assert neurons * 6 <= w17 <= neurons * 6 + 16384                # 2-byte words (16 live steps) + 4-byte trains
    assert neurons * (4 + 4 + 16) <= w18 <= neurons * 24 + 16384    # 4-byte words and trains + carried state
    assert lib.snn_rpn_head_workspace_bytes(H, W, 2, 2, 200, 8, 3) == 0 and b"multiple of 128" in lib.snn_last_error()
    assert lib.snn_rpn_head_workspace_bytes(H, W, 2, 2, 256, 33, 3) == 0
    R, K, Hd = 100, 12544, 1024
    b12 = lib.snn_box_head_workspace_bytes(R, K, Hd, 12, 3)
    b32 = lib.snn_box_head_workspace_bytes(R, K, Hd, 32, 3)
    assert R * K * 2 + 2 * R * Hd * 2 <= b12 <= R * K * 2 + 2 * R * Hd * 2 + 16384
    assert R * K * 4 + 2 * R * Hd * 4 + R * Hd * 16 <= b32 <= R * K * 4 + 2 * R * Hd * 4 + R * Hd * 16 + 16384
    assert lib.snn_box_head_workspace_bytes(R, K, Hd, 2, 3) > 0          # T < 3 runs (zero membranes, as the reference)
    assert lib.snn_box_head_workspace_bytes(R, K, Hd, 0, 3) == 0 and lib.snn_box_head_workspace_bytes(R, K, Hd, 33, 3) == 0


def test_post_processing_entry_points_validate_on_the_host(lib):
    """snn_det_postprocess / snn_rpn_nms: sizes and limits are checked before any CUDA call; the workspace query is
    host-only."""
    c = ctypes
    assert lib.snn_det_postprocess_max_candidates() == 8192
    assert lib.snn_rpn_nms_workspace_bytes(5, 2) >= 5 * 2 * 2048 * 8 and lib.snn_rpn_nms_workspace_bytes(0, 2) == 0
    one, big = (c.c_int * 1)(1000), (c.c_int * 1)(1025)
    hw = (c.c_int * 1)(768)
    rc = lib.snn_det_postprocess(None, None, one, hw, hw, 0, 9, 0.4, 0.5, 0.01, 100, 1100, None, None, None, None, None, None)
    assert rc != 0 and b"null" in lib.snn_last_error()
    cnt = (c.c_int * 2)()
    rc = lib.snn_det_postprocess(None, None, big, hw, hw, 1, 9, 0.4, 0.5, 0.01, 100, 1200, None, None, None, None,
                                 c.cast(cnt, c.c_void_p), None)
    assert rc != 0 and b"candidates" in lib.snn_last_error()           # 1025 RoIs x 8 classes > 8192
    rc = lib.snn_det_postprocess(None, None, one, hw, hw, 1, 9, 0.4, 0.5, 0.01, 100, 500, None, None, None, None,
                                 c.cast(cnt, c.c_void_p), None)
    assert rc != 0 and b"cap" in lib.snn_last_error()
    buf = c.create_string_buffer(64)
    p = c.cast(buf, c.c_void_p)
    sizes = (c.c_int * 2)(3000, 10)
    rc = lib.snn_rpn_nms(p, p, sizes, hw, hw, 2, 1, 1e-3, 0.0, 0.7, 1000, p, p, p, p, 64, None)
    assert rc != 0 and b"limit" in lib.snn_last_error()                # 3000 entries in a level > 2048
    sizes = (c.c_int * 5)(2000, 2000, 2000, 2000, 2000)
    rc = lib.snn_rpn_nms(p, p, sizes, hw, hw, 5, 1, 1e-3, 0.0, 0.7, 2000, p, p, p, p, 64, None)
    assert rc != 0 and b"keepers" in lib.snn_last_error()              # 10000 possible keepers > 8192
