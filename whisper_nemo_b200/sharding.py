"""Multi-GPU sharding of the diarization path (one process per GPU, torch.distributed).

Two levels, matching SURVEY.md section 8(e):
  * a batch of recordings is independent work: `assign_recordings` deals recordings to ranks (longest first), each
    rank diarizes its own and writes its own RTTMs; there is no data-path collective.
  * one long recording: the windows of every scale are independent (BatchNorm in eval mode, SE / pooling are
    per-window), so each rank embeds a contiguous slice and the [N_s, 192] embeddings are all-gathered
    (`all_gather_rows`, NCCL over NVLink on GPUs, gloo in the CPU tests); <= 128 MB for 4 hours of audio.
The reference has no distributed code at all; nothing here mirrors a reference interface.
"""
from typing import List, Sequence, Tuple

import torch


def is_distributed() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


def rank_world() -> Tuple[int, int]:
    if is_distributed():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def assign_recordings(durations: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first: recording indices per rank, every index exactly once."""
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for idx in sorted(range(len(durations)), key=lambda i: (-durations[i], i)):
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(idx)
        loads[r] += durations[idx]
    return [sorted(x) for x in out]


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate the row shards produced under `shard_range` into the full [n_total, d] tensor on every rank."""
    rank, world = rank_world()
    if world == 1:
        return local
    base, extra = divmod(n_total, world)
    pad_rows = base + (1 if extra else 0)
    d = local.shape[1]
    buf = torch.zeros(pad_rows, d, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * pad_rows, d, dtype=local.dtype, device=local.device)
    torch.distributed.all_gather_into_tensor(out, buf)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        pieces.append(out[r * pad_rows : r * pad_rows + (hi - lo)])
    return torch.cat(pieces, dim=0)
