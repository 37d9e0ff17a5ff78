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


class _Gathered:
    """Handle of an in-flight gather_rows: ``result()`` waits for the collective (on the current stream for NCCL) and
    returns the (B_global, n) rows in image order."""

    def __init__(self, flat, work, sizes, width, tail):
        self._flat, self._work, self._sizes, self._width, self._tail = flat, work, sizes, width, tail

    def result(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        out = self._flat.view((len(self._sizes), self._width) + self._tail)
        if all(b - a == self._width for a, b in self._sizes):
            return out.reshape((-1,) + self._tail)
        return torch.cat([out[r, :b - a] for r, (a, b) in enumerate(self._sizes)], dim=0)


def gather_rows(local_rows, num_images=None, group=None, async_op=False):
    """all_gather of per-image rows (B_local, n) -> (B_global, n) in image order on every rank (one collective kernel,
    ``all_gather_into_tensor``).  Handles ragged shards (B_global not divisible by the world size) by padding to the
    largest shard.  ``async_op=True`` returns a handle whose ``result()`` waits for the collective, so the next step's
    kernels need not queue behind the cross-rank rendezvous."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _Gathered(local_rows, None, [(0, local_rows.shape[0])], local_rows.shape[0], tuple(local_rows.shape[1:])) if async_op else local_rows
    world = dist.get_world_size(group)
    if num_images is None:
        n = torch.tensor([local_rows.shape[0]], device=local_rows.device)
        dist.all_reduce(n, group=group)
        num_images = int(n.item())
    sizes = [shard_range(num_images, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    tail = tuple(local_rows.shape[1:])
    if local_rows.shape[0] == width:
        pad = local_rows.contiguous()
    else:
        pad = local_rows.new_zeros((width,) + tail)
        pad[:local_rows.shape[0]] = local_rows
    flat = torch.empty((world * width,) + tail, dtype=pad.dtype, device=pad.device)
    work = dist.all_gather_into_tensor(flat, pad, group=group, async_op=async_op)
    h = _Gathered(flat, work if async_op else None, sizes, width, tail)
    return h if async_op else h.result()


def sample_diversity_rows(joints, batch, num_samples):
    """Per-image metric row used by bench.py: mean over joints of the across-sample std of joint positions
    (the 'sample diversity' family of metrics/eval_metrics_tracker.py, reduced on the device).  joints (B*N,J,3)."""
    return joints.view(batch, num_samples, -1, 3).std(dim=1).norm(dim=-1).mean(dim=-1, keepdim=True)
