"""Restatement of norse/torch/module/lif.py::LIFCell (0.0.7, feed-forward cell)
on top of module/snn.py::SNNCell.forward:  state = state or initial_state(x)."""
import torch

from ..functional.lif import LIFParameters, LIFFeedForwardState, lif_feed_forward_step


class LIFCell(torch.nn.Module):
    def __init__(self, p: LIFParameters = LIFParameters(), dt: float = 0.001, **kwargs):
        super().__init__()
        self.p = p          # a NamedTuple attribute: contributes no state_dict keys
        self.dt = dt

    def initial_state(self, input_tensor: torch.Tensor) -> LIFFeedForwardState:
        return LIFFeedForwardState(
            v=torch.full(input_tensor.shape, torch.as_tensor(self.p.v_leak).detach().item(),
                         device=input_tensor.device, dtype=input_tensor.dtype),
            i=torch.zeros(*input_tensor.shape, device=input_tensor.device, dtype=input_tensor.dtype),
        )

    def forward(self, input_tensor: torch.Tensor, state=None):
        state = state if state is not None else self.initial_state(input_tensor)
        return lif_feed_forward_step(input_tensor, state, self.p, self.dt)
