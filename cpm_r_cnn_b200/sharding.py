"""Image sharding of the op path across the GPUs of one box.

The ops never exchange data: every RoI reads only its own image's pyramid (`roi_batch_ind`, ROIAlign_cuda.cu:200,219),
NMS is per image / class / level, decode is per RoI.  So the path is partitioned by image, one process per GPU, and the
only collective is the one that reduces the *timing* of a measured region (and, in the surrounding training step, DDP's
gradient all-reduce, which is not part of this package).

  * image_range()    contiguous ranges per rank: the reference's inference split (pet/utils/subprocess.py:31-36,
                     `np.array_split(range(total), num_gpus)`: the first `total % world` ranks get one extra image).
  * shard_rois()     the rows of a (K,5) RoI table that belong to a rank, with the batch index rebased to the rank's
                     local image numbering (what the rank's Pooler sees: poolers.py:90-101 numbers images from 0).
  * aggregate()      whole-job throughput of a region timed on every rank: sum of units / max of time.
"""
import torch


def image_range(num_images, world_size, rank):
    """[start, end) of the images rank `rank` owns; same split as np.array_split(range(num_images), world_size)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("rank %r outside world of size %r" % (rank, world_size))
    base, extra = divmod(int(num_images), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rois(rois, num_images, world_size, rank):
    """rois (K,5) [global image index, x1, y1, x2, y2] -> (the rank's rows with local image indices, their positions)."""
    start, end = image_range(num_images, world_size, rank)
    img = rois[:, 0].long()
    sel = torch.nonzero((img >= start) & (img < end)).squeeze(1)
    local = rois[sel].clone()
    local[:, 0] -= start
    return local, sel


def aggregate(units, seconds, group=None):
    """(units, seconds) measured on this rank -> (total units, max seconds, units / s of the whole job).

    `seconds` must be a device-side duration (CUDA events) of a region bracketed by a barrier on both sides; the max
    over ranks is the job's time.  Works on any initialised backend (nccl on the box, gloo in the CPU tests) and without
    a process group (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(units), float(seconds), float(units) / float(seconds)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item()), float(t.item()), float(u.item()) / float(t.item())
