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


class RecordGather:
    """Per-step gather of the packed result records, off the critical path.

    `launch(rec)` starts an asynchronous all-gather of this rank's records (NCCL runs it on its own stream, ordered
    after the work already enqueued on the current stream) into one of two rotating output buffers and returns at
    once; `collect()` makes the current stream wait for the gather launched ONE call earlier and returns its output
    [world * J_local, K].  Called once per step this hides the exchange (a 24-byte record per window job, latency
    bound) behind the next step's ROI sampling instead of adding it to every step.  `flush()` waits for the last one.
    With world size 1 (or no process group) the records are passed through.
    """

    def __init__(self):
        self._out = [None, None]
        self._work = [None, None]
        self._k = 0

    def launch(self, rec: torch.Tensor):
        k = self._k & 1
        self._k += 1
        if not dist.is_initialized() or dist.get_world_size() == 1:
            self._out[k], self._work[k] = rec, None
            return
        world = dist.get_world_size()
        shape = (world * rec.shape[0], rec.shape[1])
        if self._out[k] is None or tuple(self._out[k].shape) != shape or self._out[k].dtype != rec.dtype:
            self._out[k] = torch.empty(shape, dtype=rec.dtype, device=rec.device)
        self._work[k] = dist.all_gather_into_tensor(self._out[k], rec.contiguous(), async_op=True)

    def _wait(self, k):
        if self._work[k] is not None:
            self._work[k].wait()          # stream-ordered for NCCL: the current stream waits, the host does not
            self._work[k] = None
        return self._out[k]

    def collect(self):
        """Output of the gather launched one `launch` ago (None before the second launch)."""
        if self._k < 2:
            return None
        return self._wait(self._k & 1)

    def flush(self):
        """Output of the most recent gather."""
        if self._k == 0:
            return None
        other = self._k & 1
        self._wait(other)
        return self._wait((self._k - 1) & 1)


def _parse_cpulist(text: str) -> set[int]:
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process to the CPUs that are local to GPU `local_rank` (sysfs local_cpulist of its PCI function), so that
    the pinned host buffers it allocates afterwards are first-touched on that NUMA node and its launch thread stays
    next to the GPU.  Returns what it found / did; a box with one NUMA node is left as it is."""
    info = {'bound': False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        base = f'/sys/bus/pci/devices/{bdf}'
        info['pci'] = bdf
        info['numa_node'] = int(open(f'{base}/numa_node').read().strip())
        cpus = _parse_cpulist(open(f'{base}/local_cpulist').read())
        allowed = os.sched_getaffinity(0)
        target = (cpus & allowed) or allowed
        info['cpus'] = len(target)
        if target != allowed:
            os.sched_setaffinity(0, target)
            info['bound'] = True
    except (OSError, AttributeError, ValueError) as e:
        info['error'] = f'{type(e).__name__}: {e}'
    return info
