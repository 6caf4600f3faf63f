"""Multi-GPU plumbing for the spiking heads: images (and their RoIs) are sharded across ranks,
one process per GPU, with NO collective on the hot path -- every image's RPN head and every
RoI's box head is independent (the reference's own multi-GPU mode is DistributedSampler + DDP,
train.py:594-601,708).  The only exchange is an end-of-batch all-gather of fixed-size per-image
records (spike-rate statistics / detection summaries), replacing the reference's
all_gather_object of pickled eval arrays (coco_eval.py:158-160, utils.py:78-91)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    lo, hi = shard_range(n_images, rank, world)
    return list(range(lo, hi))


class PendingGather:
    """Handle of an in-flight record gather (see gather_records_async)."""

    def __init__(self, work, out, counts, feat_shape, world, mx):
        self.work, self.out, self.counts, self.F, self.world, self.mx = work, out, counts, feat_shape, world, mx

    def result(self) -> torch.Tensor:
        """Make the current stream wait for the collective and return [sum(counts), F] in global image order."""
        if self.work is not None:
            self.work.wait()
        if self.world == 1:
            return self.out
        out = self.out.view(self.world, self.mx, *self.F)
        return torch.cat([out[r, : self.counts[r]] for r in range(self.world)], dim=0)


def gather_records_async(local: torch.Tensor, counts: Sequence[int]) -> PendingGather:
    """Start the all-gather of per-image records and return immediately: the collective runs on NCCL's own stream
    behind the kernels that produced `local`, and the caller's stream is NOT made to wait for it until
    `.result()` -- so the next batch's head kernels overlap the exchange (nothing on the hot path waits on a
    collective).  `local` is [n_local, F] on this rank, counts[r] the number of images rank r owns; ragged shards
    are padded to max(counts) so one fixed-size all_gather_into_tensor is used."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return PendingGather(None, local, list(counts), tuple(local.shape[1:]), 1, local.shape[0])
    world = dist.get_world_size()
    assert len(counts) == world and local.shape[0] == counts[dist.get_rank()]
    mx = max(counts)
    F = tuple(local.shape[1:])
    if local.shape[0] == mx:
        padded = local.contiguous()
    else:
        padded = local.new_zeros((mx,) + F)
        padded[: local.shape[0]] = local
    out = local.new_empty((world * mx,) + F)
    work = dist.all_gather_into_tensor(out, padded, async_op=True)
    return PendingGather(work, out, list(counts), F, world, mx)


def gather_records(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gather per-image records (blocking form of gather_records_async).  Returns [sum(counts), F] in global
    image order on every rank."""
    return gather_records_async(local, counts).result()


_DENOM_CACHE = {}


def spike_rate_records(rpn_counts: torch.Tensor, level_sizes: Sequence[Tuple[int, int]], channels: int, T_rpn: int,
                       box_counts: torch.Tensor, rois_per_image: int, hidden: int, T_det: int) -> torch.Tensor:
    """Per-image record [n_local, levels + 2] of mean spike rates: shared_lif per FPN level, then lif6
    and lif7 averaged over the image's RoIs (the quantities train.py:482,491 reads from the reference's
    spike-rate list, indices {0,3,6,9,12} and {15,16}).  Three small device ops, no host sync."""
    L, N = rpn_counts.shape
    key = (str(rpn_counts.device), tuple(level_sizes), channels, T_rpn)
    denom = _DENOM_CACHE.get(key)
    if denom is None:        # built once: torch.tensor(list, device=cuda) is a synchronous host->device copy
        denom = torch.tensor([float(h * w * channels * T_rpn) for (h, w) in level_sizes], dtype=torch.float64)
        denom = _DENOM_CACHE[key] = denom.to(rpn_counts.device)
    rpn_rates = rpn_counts.double().t() / denom                                        # [N, L]
    box_rates = box_counts.double().view(2, N, rois_per_image).sum(dim=2).t() / float(rois_per_image * hidden * T_det)
    return torch.cat([rpn_rates, box_rates], dim=1)
