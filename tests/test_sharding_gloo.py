"""Multi-GPU host logic on CPU: the path shards by image with no data-path collective; the only
collective reduces the timing.  world_size-2 gloo process groups exercise exactly the code bench.py runs under nccl."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpm_r_cnn_b200 import sharding, synthetic


def test_image_range_matches_the_reference_split():
    """pet/utils/subprocess.py:31-36 splits with np.array_split(range(total), num_gpus)."""
    for total in (0, 1, 2, 5, 16, 17, 5000):
        for world in (1, 2, 3, 4, 8):
            parts = np.array_split(range(total), world)
            covered = []
            for rank in range(world):
                s, e = sharding.image_range(total, world, rank)
                assert e - s == len(parts[rank])
                if len(parts[rank]):
                    assert s == parts[rank][0] and e == parts[rank][-1] + 1
                covered += list(range(s, e))
            assert covered == list(range(total))
    with pytest.raises(ValueError):
        sharding.image_range(4, 2, 2)


def test_shard_rois_partitions_and_rebases():
    gen = torch.Generator().manual_seed(0)
    rois = synthetic.coco_like_rois(gen, 7, 5)
    rois = rois[torch.randperm(rois.shape[0], generator=gen)]          # any row order
    seen = torch.zeros(rois.shape[0], dtype=torch.bool)
    for rank in range(2):
        local, sel = sharding.shard_rois(rois, 5, 2, rank)
        s, e = sharding.image_range(5, 2, rank)
        assert not seen[sel].any()
        seen[sel] = True
        assert torch.equal(local[:, 1:], rois[sel][:, 1:])
        assert torch.equal(local[:, 0] + s, rois[sel][:, 0])
        assert local.shape[0] == 7 * (e - s) and (local[:, 0] >= 0).all() and (local[:, 0] < e - s).all()
    assert seen.all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank builds the same global problem from the same seed and keeps its own images: no exchange needed
        gen = torch.Generator().manual_seed(1234)
        n_img = 5
        rois = synthetic.coco_like_rois(gen, 6, n_img)
        local, sel = sharding.shard_rois(rois, n_img, world, rank)
        # a rank's result depends only on its own rows (stand-in for the per-RoI op: a row-wise function)
        res = torch.zeros(rois.shape[0], dtype=torch.float64)
        res[sel] = (local[:, 1:].double() ** 2).sum(1) + local[:, 0].double() + sharding.image_range(n_img, world, rank)[0]
        dist.all_reduce(res)                                             # test-only gather of the disjoint pieces
        full = (rois[:, 1:].double() ** 2).sum(1) + rois[:, 0].double()
        # timing aggregation: units add up, the slowest rank sets the time
        units, secs, rate = sharding.aggregate(local.shape[0], 0.25 * (rank + 1))
        dist.barrier()
        if rank == 0:
            out.put((bool(torch.allclose(res, full)), units, secs, rate))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    ok, units, secs, rate = out.get()
    assert ok
    assert units == 30.0 and secs == 0.5 and rate == pytest.approx(60.0)


def test_aggregate_without_process_group():
    assert sharding.aggregate(10, 2.0) == (10.0, 2.0, 5.0)
