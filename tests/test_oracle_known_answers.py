"""Pins the CPU restatement of the Norse 0.0.7 primitives (oracle/) with
closed-form known answers -- the reference ships no
tests for this path, so these are the only independent pins."""
import math

import numpy as np
import torch

from oracle import snn_oracle as O


def test_dt_tau_products_round_to_fp32_tenth_and_fifth():
    assert O._K_MEM.dtype == torch.float32 and O._K_SYN.dtype == torch.float32
    assert O._K_MEM.item() == np.float32(0.1).item()
    assert O._K_SYN.item() == np.float32(0.2).item()


def test_encoder_constant_current_closed_form():
    # before the first spike v_n = x (1 - 0.9^n); first spike at the smallest n with v_n > 0.25
    for x in [0.1, 0.3, 0.44, 0.6, 1.0, 3.0, 30.0]:
        xs = torch.tensor([x])
        zs = O.encoder_spikes(xs, 32)
        fired = [t for t, z in enumerate(zs) if z.item() == 1.0]
        n_first = next((n for n in range(1, 33) if x * (1 - 0.9 ** n) > 0.25 + 1e-6), None)
        if n_first is None:
            assert fired == [] or fired[0] >= 31
        else:
            assert fired[0] == n_first - 1
            period = n_first                    # reset to exactly 0 -> periodic train
            assert fired[:3] == [n_first - 1 + k * period for k in range(3) if n_first - 1 + k * period < 32][:3]


def test_encoder_never_fires_below_threshold_current():
    zs = O.encoder_spikes(torch.tensor([0.2499, -5.0, 0.0]), 64)
    assert sum(float(z.sum()) for z in zs) == 0.0
    # x (1 - 0.9^8) > 0.25 <=> x > 0.439 for T = 8
    assert sum(float(z.sum()) for z in O.encoder_spikes(torch.tensor([0.43]), 8)) == 0.0
    assert sum(float(z.sum()) for z in O.encoder_spikes(torch.tensor([0.45]), 8)) == 1.0


def test_encoder_reset_is_exact_zero():
    x = torch.tensor([5.0]); v = torch.zeros(1)
    z, v = O.encoder_step(x, v)
    assert z.item() == 1.0 and v.item() == 0.0


def test_lif_one_step_delay_and_threshold():
    v = torch.zeros(4); i = torch.zeros(4)
    cur0 = torch.tensor([0.5, 1.0, 1.0001, 2.0])
    z0, v, i, _ = O.lif_step(cur0, v, i)
    assert z0.sum() == 0                      # z_0 == 0 whatever the input
    z1, v, i, vdec = O.lif_step(torch.zeros(4), v, i)
    # v_dec = 0.1 * cur0 ; spike iff 0.1*cur0 - 0.1 > 0 (strict)
    assert z1.tolist() == [0.0, 0.0, 1.0, 1.0]
    assert torch.equal(vdec, torch.tensor(0.1) * cur0)
    assert v[2].item() == 0.0 and v[3].item() == 0.0      # reset to v_reset = 0
    assert torch.allclose(i, cur0 - torch.tensor(0.2) * cur0)


def test_li_impulse_response_is_kappa():
    T = 16
    kap = O.li_kernel(T)
    assert abs(kap[0].item() - 0.1) < 1e-15 and abs(kap[1].item() - 0.17) < 1e-15
    v = torch.zeros(1); i = torch.zeros(1)
    got = []
    for t in range(T):
        v, i = O.li_step(torch.tensor([1.0 if t == 0 else 0.0]), v, i)
        got.append(v.item())
    assert np.allclose(got, kap.numpy(), rtol=0, atol=3e-7)


def test_li_linearity_closed_form_vs_stepping():
    torch.manual_seed(0)
    T = 12
    cur = torch.randn(T, 50)
    v = torch.zeros(50); i = torch.zeros(50)
    for t in range(T):
        v, i = O.li_step(cur[t], v, i)
    kap = O.li_kernel(T).to(torch.float32)
    closed = sum(kap[T - 1 - t] * cur[t] for t in range(T))
    assert torch.allclose(v, closed, rtol=0, atol=2e-6)


def test_flat_restatement_matches_stepwise_rpn():
    torch.manual_seed(1)
    w = {"s": 0.05 * torch.randn(32, 32, 3, 3), "c": 0.05 * torch.randn(3, 32, 1, 1), "b": 0.05 * torch.randn(12, 32, 1, 1)}
    feats = [2 * torch.randn(2, 32, 7, 9), 2 * torch.randn(2, 32, 4, 5)]
    for T in (1, 2, 5, 8):
        lo, bb, tr = O.rpn_head_forward(feats, w["s"], w["c"], w["b"], T, record=True)
        lo2, bb2, spk2 = O.rpn_head_flat(feats, w["s"], w["c"], w["b"], T)
        for l in range(2):
            if T > 1:
                assert torch.equal(tr[l]["spk"], spk2[l])          # dead-step identities are exact
            assert torch.allclose(lo[l], lo2[l], rtol=0, atol=5e-6)
            assert torch.allclose(bb[l], bb2[l], rtol=0, atol=5e-6)


def test_flat_restatement_matches_stepwise_box():
    torch.manual_seed(2)
    w6 = 0.1 * torch.randn(64, 80); w7 = 0.2 * torch.randn(64, 64)
    wc = 0.2 * torch.randn(5, 64); wb = 0.2 * torch.randn(20, 64)
    x = 2 * torch.randn(11, 5, 4, 4)
    for T in (3, 4, 8, 12):
        c, b, tr = O.box_head_forward(x, w6, w7, wc, wb, T, record=True)
        c2, b2, s6, s7 = O.box_head_flat(x, w6, w7, wc, wb, T)
        # spk6 at the last step is dead work the flat form never computes
        assert torch.equal(tr["spk6"][: T - 1], s6[: T - 1])
        assert torch.equal(tr["spk7"], s7)
        assert torch.allclose(c, c2, rtol=0, atol=5e-6) and torch.allclose(b, b2, rtol=0, atol=5e-6)
        assert T < 8 or tr["spk7"].sum() > 0


def test_pack_trains_roundtrip():
    spk = (torch.rand(9, 4, 5) > 0.5).to(torch.uint8)
    w = O.pack_trains(spk)
    for t in range(9):
        assert torch.equal(((w >> t) & 1).to(torch.uint8), spk[t])


def test_rate_variants_are_consistent_with_forward():
    torch.manual_seed(3)
    ws, wc, wb = 0.05 * torch.randn(16, 16, 3, 3), 0.1 * torch.randn(3, 16, 1, 1), 0.1 * torch.randn(12, 16, 1, 1)
    feats = [3 * torch.randn(2, 16, 6, 6)]
    T = 6
    rates = O.rpn_head_rates(feats, ws, wc, wb, T, 3)
    _, _, tr = O.rpn_head_forward(feats, ws, wc, wb, T, record=True)
    want = tr[0]["spk"].float().sum(0).flatten(1).div(T).mean(1)
    assert torch.allclose(rates[0][:, 0], want, atol=1e-7)
    assert rates[0][0, 1].item() == 9 * 36 * 16 * 16
    assert rates[1][0, 1].item() == 36 * 16 * 3 * 4 and rates[2][0, 1].item() == 36 * 16 * 3   # swapped, as in the reference


def test_two_fp16_pieces_of_the_row_scaled_weight_are_within_one_fp32_ulp():
    # the split the fp16x2 mode feeds the tensor cores with (csrc/aux_kernels.cuh prep_weights_kernel)
    from tests._util_cpu import split_reconstruct
    g = torch.Generator().manual_seed(3)
    for w in (torch.randn(64, 2304, generator=g) * 0.01,                       # rpn.py:78-82 init
              (torch.rand(32, 12544, generator=g) * 2 - 1) / 12544 ** 0.5):    # nn.Linear default init (fc6)
        eff = split_reconstruct(w, 2, fp16=True)
        # one fp32 ulp of the weight, or (weights below 2^-16 of the row maximum, whose pieces reach the
        # fp16 subnormals) 2^-39 of the row maximum -- far below the fp32 rounding of the row's sum
        bound = torch.maximum(w.abs().double() * 2.0 ** -23, w.abs().amax(dim=1, keepdim=True).double() * 2.0 ** -39)
        assert ((eff - w.double()).abs() <= bound).all()
        assert (eff == w.double()).float().mean() > 0.4                          # about half are exact
        eff3 = split_reconstruct(w, 3, fp16=False)
        assert torch.equal(eff3, w.double())                                     # 3 bf16 pieces: exact
