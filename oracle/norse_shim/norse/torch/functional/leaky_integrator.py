"""Restatement of norse/torch/functional/leaky_integrator.py (0.0.7).
Call sites in the reference: rpn.py:71,75,111,115; faster_rcnn.py:456,468,506,510."""
from typing import NamedTuple

import torch


class LIParameters(NamedTuple):
    tau_syn_inv: torch.Tensor = torch.as_tensor(1.0 / 5e-3)
    tau_mem_inv: torch.Tensor = torch.as_tensor(1.0 / 1e-2)
    v_leak: torch.Tensor = torch.as_tensor(0.0)


class LIState(NamedTuple):
    v: torch.Tensor
    i: torch.Tensor


def li_feed_forward_step(input_tensor, state, p: LIParameters = LIParameters(), dt: float = 0.001):
    # the input jumps the current FIRST, then the membrane integrates it
    i_new = state.i + input_tensor
    dv = dt * p.tau_mem_inv * ((p.v_leak - state.v) + i_new)
    v_new = state.v + dv
    di = -dt * p.tau_syn_inv * i_new
    i_decayed = i_new + di
    return v_new, LIState(v_new, i_decayed)
