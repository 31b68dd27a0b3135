"""Multi-GPU plumbing: frames are independent, so ranks shard the batch with no data-path
collective; one gather of the packed fixed-capacity keypoint payload per batch follows
(SURVEY.md §8e).  Backend: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous batch split: rank r owns frames [lo, hi).  Remainder frames go to the first ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_packed(packed, dst=0, group=None):
    """Gather every rank's packed result [n_r, width] on ``dst``.  Shards may differ by one frame, so
    payloads are padded to the largest shard; returns the list of per-rank tensors (trimmed) on dst,
    None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    nmax = int(max(int(s.item()) for s in sizes))
    if packed.shape[0] < nmax:
        pad = torch.zeros((nmax - packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
        packed = torch.cat([packed, pad], 0)
    out = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed.contiguous(), out, dst=dst, group=group)
    if rank != dst:
        return None
    return [o[: int(s.item())] for o, s in zip(out, sizes)]
