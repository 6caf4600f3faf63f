"""CPU-only helpers shared by the oracle tests and the GPU parity tests."""
import torch


def split_reconstruct(w, pieces, fp16=False):
    """Value (float64) the tensor cores effectively see for `pieces` 16-bit pieces (hi+mid+lo) of the
    fp32 weight matrix w [O, K].  fp16 pieces are taken from the row scaled by 2^S with
    max|row| * 2^S in [2^14, 2^15) (csrc/aux_kernels.cuh prep_weights_kernel)."""
    w = w.clone().float()
    if fp16:
        m = w.abs().amax(dim=1, keepdim=True)
        e = torch.frexp(m)[1]
        S = torch.where(m > 0, 15 - e, torch.zeros_like(e)).clamp(-100, 100)
        r = torch.ldexp(w, S)
    else:
        r = w
    tot = torch.zeros_like(r, dtype=torch.float64)
    for _ in range(pieces):
        b = r.to(torch.float16 if fp16 else torch.bfloat16).float()
        tot += b.double()
        r = r - b
    if fp16:
        tot = torch.ldexp(tot, -S)
    return tot
