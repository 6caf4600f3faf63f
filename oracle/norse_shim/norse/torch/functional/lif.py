"""Restatement of norse/torch/functional/lif.py (0.0.7): LIFParameters,
lif_feed_forward_step, lif_current_encoder.  Call sites in the reference:
rpn.py:58,67,101,106; faster_rcnn.py:444,449,452,494,499,501.

Every product/sum below is a separate fp32 torch op in exactly this order --
the CUDA kernels reproduce the same order with __fmul_rn/__fadd_rn (no FMA
contraction).  dt * tau_*_inv is evaluated first (python float times 0-dim
fp32 tensor), which rounds to 0.1f / 0.2f for dt = 0.001.
"""
from typing import NamedTuple

import torch

from .threshold import threshold


class LIFParameters(NamedTuple):
    tau_syn_inv: torch.Tensor = torch.as_tensor(1.0 / 5e-3)
    tau_mem_inv: torch.Tensor = torch.as_tensor(1.0 / 1e-2)
    v_leak: torch.Tensor = torch.as_tensor(0.0)
    v_th: torch.Tensor = torch.as_tensor(1.0)
    v_reset: torch.Tensor = torch.as_tensor(0.0)
    method: str = "super"
    alpha: float = torch.as_tensor(100.0)


class LIFFeedForwardState(NamedTuple):
    v: torch.Tensor
    i: torch.Tensor


def lif_feed_forward_step(input_tensor, state, p: LIFParameters = LIFParameters(), dt: float = 0.001):
    # voltage update from the OLD synaptic current: the input of this step
    # reaches v only at the next step (one-step delay, z_0 == 0).
    dv = dt * p.tau_mem_inv * ((p.v_leak - state.v) + state.i)
    v_decayed = state.v + dv
    di = -dt * p.tau_syn_inv * state.i
    i_decayed = state.i + di
    z_new = threshold(v_decayed - p.v_th, p.method, p.alpha)
    v_new = (1 - z_new) * v_decayed + z_new * p.v_reset
    i_new = i_decayed + input_tensor
    return z_new, LIFFeedForwardState(v=v_new, i=i_new)


def lif_current_encoder(input_current, voltage, p: LIFParameters = LIFParameters(), dt: float = 0.001):
    dv = dt * p.tau_mem_inv * ((p.v_leak - voltage) + input_current)
    voltage = voltage + dv
    z = threshold(voltage - p.v_th, p.method, p.alpha)
    voltage = voltage - z * (voltage - p.v_reset)
    return z, voltage
