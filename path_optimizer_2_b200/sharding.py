"""Data-parallel sharding of a batch of independent path QPs over the GPUs of one node.

Every instance is independent, so the batch is cut into contiguous blocks, one
per rank, with no data-path collective. The only exchange is one all-gather of the
per-instance results {cost f64, status i32, iters i32} = 16 bytes per instance, over NCCL
(NVLink 5 / NVSwitch) on the GPUs or gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `total` instances owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(cost: torch.Tensor, status: torch.Tensor, iters: torch.Tensor) -> torch.Tensor:
    """(B,) f64 cost + (B,) i32 status + (B,) i32 iters -> (B, 2) f64 (16 B per instance)."""
    si = torch.stack((status.to(torch.int32), iters.to(torch.int32)), dim=1).contiguous()
    out = torch.empty((cost.shape[0], 2), dtype=torch.float64, device=cost.device)
    out[:, 0] = cost
    out[:, 1] = si.view(torch.float64).squeeze(1)
    return out


def unpack_results(packed: torch.Tensor):
    cost = packed[:, 0].clone()
    si = packed[:, 1].contiguous().view(torch.int32).view(-1, 2)
    return cost, si[:, 0].clone(), si[:, 1].clone()


def gather_results(packed: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """All-gather equal-sized shards of packed results into (world*B, 2) on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return packed
    if out is None:
        out = torch.empty((world * packed.shape[0], 2), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out
