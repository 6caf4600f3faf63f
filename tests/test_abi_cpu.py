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
    assert lib.snn_version() == want


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
