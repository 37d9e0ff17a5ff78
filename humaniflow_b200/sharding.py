"""Multi-GPU plumbing of the hot path: shard the image axis, gather per-image rows.

The reference is single-process / single-device; this is new design (8e).  Every (image, sample)
pair is independent once the image's encoder features exist, so ranks own disjoint image ranges and run the
whole path locally.  There is NO data-path collective; the only communication is one all_gather of the
per-image metric rows (a few floats per image) per batch.  Each rank takes the matching slice of the injected
noise so the N-GPU result equals the 1-GPU result on the same global noise.
"""
import torch
import torch.distributed as dist


def shard_range(num_images, world_size, rank):
    """Contiguous [start, stop) of images owned by `rank`; the first (num_images % world_size) ranks own one more."""
    base, rem = divmod(num_images, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(t, world_size, rank, dim=0):
    """Slice of a global tensor (images, noise, targets) owned by `rank` along the image axis."""
    a, b = shard_range(t.shape[dim], world_size, rank)
    return t.narrow(dim, a, b - a)


def gather_rows(local_rows, num_images=None, group=None):
    """all_gather of per-image rows (B_local, n) -> (B_global, n) in image order on every rank.
    Handles ragged shards (B_global not divisible by the world size) by padding to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    if num_images is None:
        n = torch.tensor([local_rows.shape[0]], device=local_rows.device)
        dist.all_reduce(n, group=group)
        num_images = int(n.item())
    sizes = [shard_range(num_images, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = local_rows.new_zeros((width,) + tuple(local_rows.shape[1:]))
    pad[:local_rows.shape[0]] = local_rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def sample_diversity_rows(joints, batch, num_samples):
    """Per-image metric row used by bench.py: mean over joints of the across-sample std of joint positions
    (the 'sample diversity' family of metrics/eval_metrics_tracker.py, reduced on the device).  joints (B*N,J,3)."""
    return joints.view(batch, num_samples, -1, 3).std(dim=1).norm(dim=-1).mean(dim=-1, keepdim=True)
