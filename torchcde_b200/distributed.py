"""Batch sharding across the GPUs of one box.

Every path is independent for the fixed-step methods, so the N-GPU layout is: one process per
GPU, contiguous shards of the batch, ONE broadcast of the vector-field weights from rank 0 at
start (NCCL over NVLink; 8,448 floats at H=32, C=8), and no collective inside the time loop.
Outputs stay sharded unless ``gather=True``.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_paths, world_size, rank):
    """Contiguous, balanced split: the first ``n_paths % world_size`` ranks get one extra path."""
    base, extra = divmod(n_paths, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_field(func, src=0, group=None):
    """Make every rank hold rank ``src``'s vector-field parameters (one flat broadcast)."""
    params = [p for p in func.parameters()]
    if not params:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    offset = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[offset:offset + n].view_as(p))
            offset += n


def cdeint_sharded(solve, control_builder, func, z0_shard, t, gather=False, group=None, **kwargs):
    """Solve this rank's shard.  ``solve`` is ``torchcde_b200.cdeint`` (injected so the host logic
    can be exercised on CPU with gloo); ``control_builder()`` returns this rank's control."""
    broadcast_field(func, src=0, group=group)
    out = solve(control_builder(), func, z0_shard, t, **kwargs)
    if not gather:
        return out
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=out.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([out.size(0)], dtype=torch.int64, device=out.device), group=group)
    biggest = int(max(int(s) for s in sizes))
    padded = out.new_zeros(biggest, *out.shape[1:])
    padded[:out.size(0)] = out
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:int(s)] for p, s in zip(parts, sizes)], dim=0)
