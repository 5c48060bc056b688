"""Host-side multi-GPU plumbing: frames are independent, so a batch is sharded over ranks (one process per
GPU) with NO data-path collective.  torch.distributed is used only to agree on the timing:
aggregate frames/s = (sum of frames over ranks) / (max elapsed over ranks).
Works on any backend (nccl on the GPUs, gloo in the CPU tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a global batch owned by `rank` (sizes differ by at most one frame)."""
    if world <= 0 or not (0 <= rank < world) or total_frames < 0:
        raise ValueError(f"bad shard request: total={total_frames} rank={rank} world={world}")
    base, rem = divmod(total_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank seed of the synthetic inputs ('per-rank seed = base + rank')."""
    return int(base_seed) + int(rank)


def aggregate_throughput(local_frames: int, local_elapsed_ms: float, device=None):
    """(total frames, max elapsed ms, frames/s) over all ranks; identity when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = device if device is not None else torch.device("cpu")
        frames = torch.tensor([float(local_frames)], dtype=torch.float64, device=dev)
        elapsed = torch.tensor([float(local_elapsed_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(frames, op=dist.ReduceOp.SUM)
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        total, ms = float(frames.item()), float(elapsed.item())
    else:
        total, ms = float(local_frames), float(local_elapsed_ms)
    return total, ms, (total / (ms * 1e-3) if ms > 0 else float("nan"))
