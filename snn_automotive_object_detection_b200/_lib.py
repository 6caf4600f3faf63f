"""ctypes binding of libsnn_heads_b200.so (C ABI in include/snn_heads.h).

The product path has no CPU or eager fallback: if the CUDA library cannot be
loaded, or a call is made with non-CUDA tensors, a RuntimeError is raised.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnn_heads_b200.so")

MODE_FP32_EXACT, MODE_BF16, MODE_BF16X2, MODE_FP16X2, MODE_FP16 = 0, 1, 2, 3, 4
MODES = {"fp32_exact": MODE_FP32_EXACT, "fp32": MODE_FP32_EXACT, "bf16x3": MODE_FP32_EXACT, "bf16": MODE_BF16,
         "bf16x2": MODE_BF16X2, "fp16x2": MODE_FP16X2, "fp16": MODE_FP16}
# 16-bit pieces per weight and the torch dtype of the {0,1} spike planes each mode contracts
PIECES = {MODE_FP32_EXACT: 3, MODE_BF16: 1, MODE_BF16X2: 2, MODE_FP16X2: 2, MODE_FP16: 1}
FP16_MODES = (MODE_FP16X2, MODE_FP16)

# every symbol include/snn_heads.h declares
EXPORTS = [
    "snn_version", "snn_last_error", "snn_train_word_bytes", "snn_mode_pieces", "snn_prepared_weight_bytes",
    "snn_prepare_conv3x3_weights", "snn_prepare_fc_weights", "snn_rpn_head_workspace_bytes", "snn_rpn_head_forward",
    "snn_box_head_workspace_bytes", "snn_box_head_forward", "snn_fc_lif_layer", "snn_encode_rows",
    "snn_last_launch_count", "snn_set_cta_group", "snn_profile_enable", "snn_profile_read", "snn_rpn_decode_selected",
    "snn_roi_align_encode", "snn_box_head_forward_encoded", "snn_encoder_table", "snn_encoder_selftest", "snn_set_fc_tiling", "snn_set_role_timers",
    "snn_set_clock_probe", "snn_host_cache_stats", "snn_rpn_topk_keys", "snn_set_roi_kernel", "snn_li_readout_nhwc", "snn_li_readout_rows", "snn_rpn_topk_workspace_bytes", "snn_rpn_topk_select", "snn_set_conv_multicast",
    "snn_det_postprocess_max_candidates", "snn_det_postprocess", "snn_rpn_nms_workspace_bytes", "snn_rpn_nms",
]
# the ABI the argtypes below describe (include/snn_heads.h SNN_ABI_VERSION); a library of another version is refused
EXPECTED_ABI = 8
PHASES = ["rpn_encoder", "rpn_conv_lif_gemm", "rpn_readout", "box_encoder", "fc6_lif_gemm", "fc7_lif_gemm", "box_readout"]

_lock = threading.Lock()
_lib = None


def _declare(lib):
    c = ctypes
    vp, i, sz = c.c_void_p, c.c_int, c.c_size_t
    pi = c.POINTER(c.c_int)
    pvp = c.POINTER(c.c_void_p)
    lib.snn_version.restype = i
    lib.snn_last_error.restype = c.c_char_p
    lib.snn_train_word_bytes.argtypes = [i]; lib.snn_train_word_bytes.restype = i
    lib.snn_mode_pieces.argtypes = [i]; lib.snn_mode_pieces.restype = i
    lib.snn_prepared_weight_bytes.argtypes = [i, i, i]; lib.snn_prepared_weight_bytes.restype = sz
    lib.snn_prepare_conv3x3_weights.argtypes = [vp, i, i, i, vp, vp]; lib.snn_prepare_conv3x3_weights.restype = i
    lib.snn_prepare_fc_weights.argtypes = [vp, i, i, i, vp, vp]; lib.snn_prepare_fc_weights.restype = i
    lib.snn_rpn_head_workspace_bytes.argtypes = [pi, pi, i, i, i, i, i]; lib.snn_rpn_head_workspace_bytes.restype = sz
    lib.snn_rpn_head_forward.argtypes = [pvp, pi, pi, i, i, i, i, i, i, vp, vp, vp, pvp, pvp, pvp, vp, vp, sz, vp]
    lib.snn_rpn_head_forward.restype = i
    lib.snn_box_head_workspace_bytes.argtypes = [i, i, i, i, i]; lib.snn_box_head_workspace_bytes.restype = sz
    lib.snn_box_head_forward.argtypes = [vp, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.snn_box_head_forward.restype = i
    lib.snn_fc_lif_layer.argtypes = [vp, i, i, i, i, i, i, i, i, i, vp, vp, vp, i, vp]; lib.snn_fc_lif_layer.restype = i
    lib.snn_encode_rows.argtypes = [vp, i, i, i, vp, vp]; lib.snn_encode_rows.restype = i
    lib.snn_rpn_decode_selected.argtypes = [pvp, pvp, pvp, pi, pi, pi, pi, pi, i, i, i, vp, vp, vp, vp, vp, vp]
    lib.snn_rpn_decode_selected.restype = i
    lib.snn_rpn_topk_keys.argtypes = [pvp, pi, pi, i, i, i, pvp, vp]; lib.snn_rpn_topk_keys.restype = i
    lib.snn_rpn_topk_workspace_bytes.argtypes = [i, i]; lib.snn_rpn_topk_workspace_bytes.restype = sz
    lib.snn_rpn_topk_select.argtypes = [pvp, pi, pi, i, i, i, i, vp, vp, sz, vp]; lib.snn_rpn_topk_select.restype = i
    lib.snn_det_postprocess_max_candidates.argtypes = []; lib.snn_det_postprocess_max_candidates.restype = i
    f = c.c_float
    lib.snn_det_postprocess.argtypes = [vp, vp, pi, pi, pi, i, i, f, f, f, i, i, vp, vp, vp, vp, vp, vp]
    lib.snn_det_postprocess.restype = i
    lib.snn_rpn_nms_workspace_bytes.argtypes = [i, i]; lib.snn_rpn_nms_workspace_bytes.restype = sz
    lib.snn_rpn_nms.argtypes = [vp, vp, pi, pi, pi, i, i, f, f, f, i, vp, vp, vp, vp, sz, vp]; lib.snn_rpn_nms.restype = i
    lib.snn_box_head_forward_encoded.argtypes = lib.snn_box_head_forward.argtypes
    lib.snn_box_head_forward_encoded.restype = i
    lib.snn_roi_align_encode.argtypes = [pvp, pi, pi, c.POINTER(c.c_float), i, i, vp, vp, i, i, i, i, vp, vp, vp]
    lib.snn_roi_align_encode.restype = i
    lib.snn_encoder_table.argtypes = [c.POINTER(c.c_float), c.POINTER(c.c_uint)]; lib.snn_encoder_table.restype = None
    lib.snn_encoder_selftest.argtypes = [i, vp, vp]; lib.snn_encoder_selftest.restype = i
    lib.snn_last_launch_count.restype = i
    lib.snn_set_cta_group.argtypes = [i]; lib.snn_set_cta_group.restype = None
    lib.snn_set_fc_tiling.argtypes = [i, i, i]; lib.snn_set_fc_tiling.restype = None
    lib.snn_set_role_timers.argtypes = [vp, i]; lib.snn_set_role_timers.restype = None
    lib.snn_set_clock_probe.argtypes = [vp]; lib.snn_set_clock_probe.restype = None
    lib.snn_set_roi_kernel.argtypes = [i]; lib.snn_set_roi_kernel.restype = None
    lib.snn_set_conv_multicast.argtypes = [i]; lib.snn_set_conv_multicast.restype = None
    pd = c.POINTER(c.c_double)
    lib.snn_li_readout_nhwc.argtypes = [vp, i, i, i, i, pd, vp, i, vp, i, vp, vp, vp]; lib.snn_li_readout_nhwc.restype = i
    lib.snn_li_readout_rows.argtypes = [vp, i, i, i, pd, vp, i, vp, i, vp, vp, vp]; lib.snn_li_readout_rows.restype = i
    ull = c.POINTER(c.c_ulonglong)
    lib.snn_host_cache_stats.argtypes = [ull, ull]; lib.snn_host_cache_stats.restype = None
    lib.snn_profile_enable.argtypes = [i]; lib.snn_profile_enable.restype = None
    lib.snn_profile_read.argtypes = [c.POINTER(c.c_float), pi]; lib.snn_profile_read.restype = i


def load():
    """Load (once) and return the ctypes library.  Fails loudly when it is missing."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                    "(nvcc, sm_100a).  There is no CPU / eager fallback for the spiking heads.")
            lib = ctypes.CDLL(LIB_PATH)
            lib.snn_version.restype = ctypes.c_int
            have = lib.snn_version()
            if have != EXPECTED_ABI:
                raise RuntimeError(
                    f"{LIB_PATH} has ABI version {have}, this package binds version {EXPECTED_ABI}: the library is "
                    "stale -- rebuild it with `python __graft_entry__.py build`")
            _declare(lib)
            _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().snn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def mode_id(mode):
    if isinstance(mode, int):
        if mode not in PIECES:
            raise ValueError(f"unknown mode {mode}")
        return mode
    try:
        return MODES[str(mode).lower()]
    except KeyError:
        raise ValueError(f"unknown mode {mode!r}; expected one of {sorted(MODES)}") from None


def host_cache_stats():
    """(hits, misses) of this thread's TMA tensor-map cache."""
    h, m = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    load().snn_host_cache_stats(ctypes.byref(h), ctypes.byref(m))
    return h.value, m.value


def profile_enable(on, phases=None):
    """Per-phase CUDA-event timing inside the library: off, every phase, or only the named `phases`."""
    if on and phases:
        mask = 0
        for name in phases:
            mask |= 1 << (1 + PHASES.index(name))
        load().snn_profile_enable(mask)
    else:
        load().snn_profile_enable(1 if on else 0)


def profile_read():
    """{phase: (total_ms, n_forwards)} of the forwards since the last read."""
    ms = (ctypes.c_float * len(PHASES))()
    cnt = (ctypes.c_int * len(PHASES))()
    check(load().snn_profile_read(ms, cnt), "snn_profile_read")
    return {name: (float(ms[k]), int(cnt[k])) for k, name in enumerate(PHASES)}
