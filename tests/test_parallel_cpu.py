"""World-size-2 gloo tests (CPU) of the sharding / gather plumbing used at N > 1."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snn_automotive_object_detection_b200 import parallel as P


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = P.shard_range(n, r, world)
                assert 0 <= hi - lo <= n // world + 1
                got += list(range(lo, hi))
            assert got == list(range(n))


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = [len(P.shard_images(n_images, r, world)) for r in range(world)]
        mine = P.shard_images(n_images, rank, world)
        local = torch.tensor([[float(i), 10.0 * i, -1.0 * i] for i in mine]).view(len(mine), 3)
        full = P.gather_records(local, counts)
        # the pipelined form used by bench.py: start two exchanges, consume them later, in order
        h1 = P.gather_records_async(local, counts)
        h2 = P.gather_records_async(local * 2, counts)
        assert torch.equal(h1.result(), full) and torch.equal(h2.result(), full * 2)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5, 1])
def test_gather_records_gloo_world2(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + n_images
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), 10.0 * i, -1.0 * i] for i in range(n_images)]
    for _rank, full in res:
        assert full == want


def test_spike_rate_records_layout():
    rpn_counts = torch.tensor([[256 * 6 * 8, 0], [0, 256 * 2 * 4]])          # [levels=2, N=2]
    box_counts = torch.zeros(2, 6, dtype=torch.int32)                       # 3 RoIs per image
    box_counts[0, :3] = 16 * 12                                             # image 0: every lif6 unit always fires
    rec = P.spike_rate_records(rpn_counts, [(2, 3), (1, 1)], 256, 8, box_counts, 3, 16, 12)
    assert rec.shape == (2, 4)
    assert rec[0, 0].item() == 1.0 and rec[1, 0].item() == 0.0
    assert rec[1, 1].item() == 1.0 and rec[0, 2].item() == 1.0 and rec[1, 2].item() == 0.0
