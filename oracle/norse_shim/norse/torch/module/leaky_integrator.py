"""Restatement of norse/torch/module/leaky_integrator.py::LICell (0.0.7)."""
import torch

from ..functional.leaky_integrator import LIParameters, LIState, li_feed_forward_step


class LICell(torch.nn.Module):
    def __init__(self, p: LIParameters = LIParameters(), dt: float = 0.001, **kwargs):
        super().__init__()
        self.p = p
        self.dt = dt

    def initial_state(self, input_tensor: torch.Tensor) -> LIState:
        return LIState(
            v=torch.full(input_tensor.shape, torch.as_tensor(self.p.v_leak).detach().item(),
                         device=input_tensor.device, dtype=input_tensor.dtype),
            i=torch.zeros(*input_tensor.shape, device=input_tensor.device, dtype=input_tensor.dtype),
        )

    def forward(self, input_tensor: torch.Tensor, state=None):
        state = state if state is not None else self.initial_state(input_tensor)
        return li_feed_forward_step(input_tensor, state, self.p, self.dt)
