"""GPU building-block parity: encoder, the tcgen05 contraction (raw currents) and the fused LIF
epilogue, each against the CPU oracle, through the C ABI.  Run on the B200 box: pytest -m gpu."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import snn_oracle as O
from snn_automotive_object_detection_b200 import _lib
from tests._util import vp, stream, prepared_fc, split_reconstruct, unpack_trains, _TRAIN_DTYPE

pytestmark = pytest.mark.gpu


def test_library_reports_version_and_rejects_bad_args():
    lib = _lib.load()
    assert lib.snn_version() == _lib.EXPECTED_ABI
    rc = lib.snn_fc_lif_layer(None, 1, 0, 1, 64, 128, 8, 0, 7, 0, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.snn_last_error()


def pack_words(z, bit0, nbytes):
    """[T_live, ...] {0,1} spikes -> spike-train words (bit bit0 + t = z[t]) of `nbytes` bytes."""
    w = torch.zeros(z.shape[1:], dtype=torch.int64)
    for t in range(z.shape[0]):
        w |= z[t].to(torch.int64) << (bit0 + t)
    if nbytes == 1:
        return torch.from_numpy(w.numpy().astype(np.uint8))
    if nbytes == 2:
        return torch.from_numpy(w.numpy().astype(np.uint16).view(np.int16))
    return torch.from_numpy(w.numpy().astype(np.uint32).view(np.int32))


@pytest.mark.parametrize("T", [1, 7, 8, 12, 16, 17, 32])
def test_encoder_rows_bit_exact(T):
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(37, 192, generator=g) * 1.5)
    x[0, :8] = torch.tensor([0.25, 0.2500001, 0.439, 0.44, -1.0, 0.0, 1e-30, 100.0])
    # the ends of the lookup table's clamp, non-finite inputs (+inf spikes exactly once in Norse's arithmetic: the reset
    # inf - inf leaves NaN) and both neighbours of every one of the encoder's 32 thresholds
    x[1, :8] = torch.tensor([float("inf"), -float("inf"), float("nan"), 4.0, 3.9999998, 4.0000005, 3.0e38, 2.5])
    thr = (ctypes.c_float * 33)()
    lib.snn_encoder_table(thr, None)
    t = torch.tensor(thr[1:33], dtype=torch.float32)
    x[2, :32] = t; x[3, :32] = torch.nextafter(t, torch.zeros(32)); x[4, :32] = torch.nextafter(t, torch.full((32,), 9.0))
    xd = x.cuda()
    wb = 1 if T <= 8 else 2 if T <= 16 else 4
    z = torch.zeros(37, 192, dtype=_TRAIN_DTYPE[wb], device="cuda")
    _lib.check(lib.snn_encode_rows(vp(xd), 37, 192, T, vp(z), stream()), "encode_rows")
    torch.cuda.synchronize()
    ref = torch.stack(O.encoder_spikes(x, T))
    assert torch.equal(unpack_trains(z.cpu(), T).float(), ref)
    assert T < 8 or ref.sum() > 0


@pytest.mark.parametrize("T", [1, 3, 4, 7, 8, 10, 11, 12, 15, 16, 23, 24, 31, 32])
def test_encoder_comparator_bank_equals_simulation_for_every_fp32_input(T):
    """The comparator-bank encoder against the step-by-step lif_current_encoder simulation on the device, over ALL
    2^32 fp32 bit patterns (NaNs, infinities, subnormals, negatives included)."""
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(lib.snn_encoder_selftest(T, vp(bad), stream()), "encoder_selftest")
    torch.cuda.synchronize()
    assert bad.item() == 0


def _fc_case(R, K, M, T, t0, T_live, mode, cg, seed=0, density=0.15, in_bit0=0, in_wb=None):
    """One fc spiking layer on random input spike trains; returns the oracle-side spikes and the kernel's
    trains / raw currents."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    z = (torch.rand(T_live, R, K, generator=g) < density).float()
    w = torch.randn(M, K, generator=g) * (1.2 / np.sqrt(density * K))
    if in_wb is None:
        nb = in_bit0 + T_live
        in_wb = 1 if nb <= 8 else 2 if nb <= 16 else 4
    words = pack_words(z, in_bit0, in_wb)
    # bits outside [in_bit0, in_bit0 + T_live) must be ignored by the kernel: set them
    junk = torch.full_like(words, -1) if in_wb > 1 else torch.full_like(words, 255)
    mask = pack_words(torch.ones_like(z), in_bit0, in_wb)
    words = words | (junk & ~mask)
    zd = words.cuda()
    wd = w.cuda()
    wp = prepared_fc(wd, mode)
    tb = lib.snn_train_word_bytes(T)
    trains = torch.zeros(R, M, dtype=_TRAIN_DTYPE[tb], device="cuda")
    dump = torch.full((T_live, R, M), float("nan"), device="cuda")
    rc = lib.snn_fc_lif_layer(vp(zd), in_wb, in_bit0, R, K, M, T, t0, T_live, mode, vp(wp), vp(trains), vp(dump), cg,
                              stream())
    _lib.check(rc, "fc_lif_layer")
    torch.cuda.synchronize()
    return z, w, trains.cpu(), dump.cpu()


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mode,pieces", [(1, 1), (2, 2), (0, 3), (3, 2), (4, 1)])
def test_fc_contraction_currents(cg, mode, pieces):
    R, K, M, T, T_live = 50, 192, 256, 8, 6
    z, w, trains, dump = _fc_case(R, K, M, T, 0, T_live, mode, cg)
    w_eff = split_reconstruct(w, pieces, fp16=mode in _lib.FP16_MODES)
    ref = torch.einsum("trk,mk->trm", z.double(), w_eff.double())
    err = (dump.double() - ref).abs().max().item()
    assert not torch.isnan(dump).any(), "some accumulator columns were never written"
    assert err < 2e-5, f"max |cur - ref| = {err}"
    if mode == 3:       # two fp16 pieces of the row-scaled weight: within one fp32 ulp of every weight
        bound = torch.maximum(w.abs().double() * 2.0 ** -23, w.abs().amax(dim=1, keepdim=True).double() * 2.0 ** -39)
        assert ((w_eff - w.double()).abs() <= bound).all()
    if pieces == 3 or mode == 3:     # these splits reproduce the fp32 weights themselves (to <= 1 ulp)
        ref32 = torch.einsum("trk,mk->trm", z.double(), w.double())
        assert (dump.double() - ref32).abs().max().item() < 2e-5


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("T,t0,T_live,R,K,M", [(8, 0, 6, 50, 192, 256), (12, 1, 10, 77, 128, 512),
                                             (16, 0, 15, 33, 64, 256), (5, 0, 4, 130, 256, 256)])
@pytest.mark.parametrize("mode", [0, 3])
def test_fc_lif_epilogue_is_exact_given_currents(cg, T, t0, T_live, R, K, M, mode):
    z, w, trains, dump = _fc_case(R, K, M, T, t0, T_live, mode, cg, seed=T)
    # LIF recurrence over the kernel's own currents: must match the oracle bit for bit
    spk = O._lif_unroll(dump, T, t0=t0)                      # [T,R,M]
    got = unpack_trains(trains, T)
    assert torch.equal(got, spk)
    assert spk.sum() > 0 and spk[0].sum() == 0


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("T,t0,T_live,in_bit0,in_wb", [(8, 1, 6, 1, 1), (12, 1, 10, 1, 2), (16, 0, 15, 1, 2),
                                                      (20, 1, 18, 1, 4), (32, 0, 31, 0, 4), (12, 0, 3, 5, 4)])
def test_fc_spike_word_operand_formats(cg, T, t0, T_live, in_bit0, in_wb):
    # every input word size / bit offset the producers expand, junk bits outside the live window ignored
    R, K, M = 45, 128, 256
    z, w, trains, dump = _fc_case(R, K, M, T, t0, T_live, 0, cg, seed=T + in_bit0, in_bit0=in_bit0, in_wb=in_wb)
    ref = torch.einsum("trk,mk->trm", z.double(), w.double())
    assert not torch.isnan(dump).any()
    assert (dump.double() - ref).abs().max().item() < 2e-5
    assert torch.equal(unpack_trains(trains, T), O._lif_unroll(dump, T, t0=t0))


def test_fc_large_k_many_tiles():
    # K = 12544 (196 k-blocks, ring wrap-around many times), more tiles than SMs
    R, K, M, T, T_live = 700, 12544, 1024, 12, 10
    for cg in (1, 2):
        z, w, trains, dump = _fc_case(R, K, M, T, 0, T_live, 1, cg, seed=3, density=0.05)
        w_eff = split_reconstruct(w, 1)
        ref = torch.einsum("trk,mk->trm", z[:, :64].double(), w_eff.double())
        assert (dump[:, :64].double() - ref).abs().max().item() < 5e-5
        ref = torch.einsum("trk,mk->trm", z[:, -40:].double(), w_eff.double())
        assert (dump[:, -40:].double() - ref).abs().max().item() < 5e-5
        assert torch.equal(unpack_trains(trains, T), O._lif_unroll(dump, T))


@pytest.fixture
def dual_tiles():
    """Force dual fc tiles (2J units per tile, both accumulator buffers fed from every weight tile) wherever the
    tile shape allows it; restore the automatic choice afterwards."""
    lib = _lib.load()
    lib.snn_set_fc_tiling(2, 0, 0)
    yield lib
    lib.snn_set_fc_tiling(0, 0, 0)


@pytest.mark.parametrize("T,t0,T_live,R,K,M,mode", [
    (8, 0, 6, 50, 192, 256, 0), (12, 1, 10, 77, 128, 512, 3), (16, 0, 15, 33, 64, 256, 0), (5, 0, 4, 130, 256, 256, 3),
    (12, 0, 11, 41, 320, 256, 3),      # 11 live steps in a 12-step box (a zero padding row per unit)
    (12, 0, 11, 1, 64, 256, 0),        # a single RoI: buffer 1 and the peer CTA see only zero-filled rows
    (3, 0, 1, 300, 128, 256, 1),
])
def test_fc_dual_tiles_currents_and_spikes(dual_tiles, T, t0, T_live, R, K, M, mode):
    z, w, trains, dump = _fc_case(R, K, M, T, t0, T_live, mode, 2, seed=T + R)
    pieces = dual_tiles.snn_mode_pieces(mode)
    w_eff = split_reconstruct(w, pieces, fp16=mode in _lib.FP16_MODES)
    ref = torch.einsum("trk,mk->trm", z.double(), w_eff.double())
    assert not torch.isnan(dump).any(), "some accumulator columns were never written"
    assert (dump.double() - ref).abs().max().item() < 2e-5
    assert torch.equal(unpack_trains(trains, T), O._lif_unroll(dump, T, t0=t0))


@pytest.mark.parametrize("in_bit0,in_wb,T,t0,T_live", [(1, 1, 8, 1, 6), (1, 2, 12, 1, 10), (5, 4, 12, 0, 3)])
def test_fc_dual_tiles_word_formats(dual_tiles, in_bit0, in_wb, T, t0, T_live):
    z, w, trains, dump = _fc_case(45, 128, 256, T, t0, T_live, 0, 2, seed=T + in_bit0, in_bit0=in_bit0, in_wb=in_wb)
    ref = torch.einsum("trk,mk->trm", z.double(), w.double())
    assert not torch.isnan(dump).any()
    assert (dump.double() - ref).abs().max().item() < 2e-5
    assert torch.equal(unpack_trains(trains, T), O._lif_unroll(dump, T, t0=t0))


def test_fc_dual_tiles_are_the_default_for_long_contractions_and_match_single_tiles():
    # K = 12544: dual tiles by default; the same layer with single tiles must give the same currents and spikes
    lib = _lib.load()
    # R = 2500 x M = 512: 79 dual unit tiles x 2 = 158 tiles on 74 CTA pairs = 2 full waves + a tail of 10 tiles,
    # which the automatic choice runs as single tiles in a second launch
    R, K, M, T, T_live = 2500, 12544, 512, 12, 11
    z, w, trains_d, dump_d = _fc_case(R, K, M, T, 0, T_live, 3, 2, seed=11, density=0.05)
    assert lib.snn_last_launch_count() == 2
    lib.snn_set_fc_tiling(1, 0, 0)
    try:
        _, _, trains_s, dump_s = _fc_case(R, K, M, T, 0, T_live, 3, 2, seed=11, density=0.05)
    finally:
        lib.snn_set_fc_tiling(0, 0, 0)
    assert not torch.isnan(dump_d).any()
    assert torch.equal(dump_d, dump_s), "same k order per accumulator: the currents are bit-identical"
    assert torch.equal(trains_d, trains_s)
    w_eff = split_reconstruct(w, 2, fp16=True)
    ref = torch.einsum("trk,mk->trm", z[:, -50:].double(), w_eff.double())
    assert (dump_d[:, -50:].double() - ref).abs().max().item() < 1e-4
