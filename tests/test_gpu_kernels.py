"""GPU building-block parity: encoder, the tcgen05 contraction (raw currents) and the fused LIF
epilogue, each against the CPU oracle, through the C ABI.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from oracle import snn_oracle as O
from snn_automotive_object_detection_b200 import _lib
from tests._util import vp, stream, prepared_fc, split_reconstruct, unpack_trains, _TRAIN_DTYPE

pytestmark = pytest.mark.gpu


def test_library_reports_version_and_rejects_bad_args():
    lib = _lib.load()
    assert lib.snn_version() == 2
    rc = lib.snn_fc_lif_layer(None, 1, 64, 128, 8, 0, 7, 0, None, None, None, 0, 0, None, 0, None)
    assert rc == -1 and b"null" in lib.snn_last_error()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("T", [1, 7, 8, 12, 32])
def test_encoder_rows_bit_exact(T, mode):
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(37, 192, generator=g) * 1.5)
    x[0, :8] = torch.tensor([0.25, 0.2500001, 0.439, 0.44, -1.0, 0.0, 1e-30, 100.0])
    xd = x.cuda()
    z = torch.empty(T, 37, 192, dtype=torch.float16 if mode in _lib.FP16_MODES else torch.bfloat16, device="cuda")
    _lib.check(lib.snn_encode_rows(vp(xd), 37, 192, T, mode, vp(z), stream()), "encode_rows")
    torch.cuda.synchronize()
    ref = torch.stack(O.encoder_spikes(x, T))
    assert torch.equal(z.float().cpu(), ref)
    assert T < 8 or ref.sum() > 0


def _fc_case(R, K, M, T, t0, T_live, mode, cg, seed=0, density=0.15):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    z = (torch.rand(T_live, R, K, generator=g) < density).float()
    w = torch.randn(M, K, generator=g) * (1.2 / np.sqrt(density * K))
    sdt = torch.float16 if mode in _lib.FP16_MODES else torch.bfloat16
    zd = z.to(sdt).cuda()
    wd = w.cuda()
    wp = prepared_fc(wd, mode)
    tb = lib.snn_train_word_bytes(T)
    trains = torch.zeros(R, M, dtype=_TRAIN_DTYPE[tb], device="cuda")
    dump = torch.full((T_live, R, M), float("nan"), device="cuda")
    planes = torch.zeros(T, R, M, dtype=sdt, device="cuda")
    rc = lib.snn_fc_lif_layer(vp(zd), R, K, M, T, t0, T_live, mode, vp(wp), vp(trains), vp(planes), 0, T, vp(dump), cg,
                              stream())
    _lib.check(rc, "fc_lif_layer")
    torch.cuda.synchronize()
    return z, w, trains.cpu(), dump.cpu(), planes.float().cpu()


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mode,pieces", [(1, 1), (2, 2), (0, 3), (3, 2), (4, 1)])
def test_fc_contraction_currents(cg, mode, pieces):
    R, K, M, T, T_live = 50, 192, 256, 8, 6
    z, w, trains, dump, planes = _fc_case(R, K, M, T, 0, T_live, mode, cg)
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
    z, w, trains, dump, planes = _fc_case(R, K, M, T, t0, T_live, mode, cg, seed=T)
    # LIF recurrence over the kernel's own currents: must match the oracle bit for bit
    spk = O._lif_unroll(dump, T, t0=t0)                      # [T,R,M]
    got = unpack_trains(trains, T)
    assert torch.equal(got, spk)
    assert torch.equal(planes, spk.float())
    assert spk.sum() > 0 and spk[0].sum() == 0


def test_fc_large_k_many_tiles():
    # K = 12544 (196 k-blocks, ring wrap-around many times), more tiles than SMs
    R, K, M, T, T_live = 700, 12544, 1024, 12, 10
    for cg in (1, 2):
        z, w, trains, dump, planes = _fc_case(R, K, M, T, 0, T_live, 1, cg, seed=3, density=0.05)
        w_eff = split_reconstruct(w, 1)
        ref = torch.einsum("trk,mk->trm", z[:, :64].double(), w_eff.double())
        assert (dump[:, :64].double() - ref).abs().max().item() < 5e-5
        ref = torch.einsum("trk,mk->trm", z[:, -40:].double(), w_eff.double())
        assert (dump[:, -40:].double() - ref).abs().max().item() < 5e-5
        assert torch.equal(unpack_trains(trains, T), O._lif_unroll(dump, T))
