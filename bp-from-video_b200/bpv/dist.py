"""Multi-GPU plumbing: streams are independent (each SignalProcessor owns one SignalStore,
signal_processor.py:115; ROI pairs are formed inside a stream, :299), so the path shards by stream with
NO data-path collective.  The only exchange is a gather of the per-stream result records
(bpm[R], ptt_ms[P], peak_idx[R], lag_idx[P]) over NCCL (gloo on CPU in the tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(num_streams: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of streams owned by `rank` (all ROIs and pairs of a stream stay together)."""
    base, rem = divmod(num_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's env; initialises the process group when world > 1."""
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29577')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def gather_records(packed: torch.Tensor, counts: list[int] | None = None) -> torch.Tensor:
    """All-gather the per-job records [J_local, K] of every rank into [sum J, K], rank order = stream order.
    `counts` = J_local of every rank when they differ (ragged shards are padded to the max and trimmed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return packed
    world = dist.get_world_size()
    if counts is None:
        counts = [packed.shape[0]] * world
    jmax = max(counts)
    buf = packed
    if packed.shape[0] != jmax:
        buf = torch.zeros((jmax, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        buf[:packed.shape[0]] = packed
    out = torch.empty((world * jmax, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    if all(c == jmax for c in counts):
        return out
    return torch.cat([out[r * jmax:r * jmax + counts[r]] for r in range(world)], dim=0)
