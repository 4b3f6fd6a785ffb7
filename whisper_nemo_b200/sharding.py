"""Multi-GPU sharding of the diarization path (one process per GPU, torch.distributed).

Two levels, matching SURVEY.md section 8(e):
  * a batch of recordings is independent work: `assign_recordings` deals recordings to ranks (longest first), each
    rank diarizes its own and writes its own RTTMs; there is no data-path collective.
  * one long recording: the windows of every scale are independent (BatchNorm in eval mode, SE / pooling are
    per-window), so each rank embeds a contiguous slice and the [N_s, 192] embeddings are all-gathered
    (`all_gather_rows`, NCCL over NVLink on GPUs, gloo in the CPU tests); <= 128 MB for 4 hours of audio.
The reference has no distributed code at all; nothing here mirrors a reference interface.
"""
from typing import Dict, List, Sequence, Tuple

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


def round_robin_counts_and_order(n_items: int, world: int) -> Tuple[List[int], List[int]]:
    """Items dealt round-robin (item i to rank i % world) and gathered in rank order: (items per rank, the original index of
    every gathered row) -- `out[order[g]] = gathered[g]` restores the original order (the NME sweep's p values)."""
    counts = [len(range(r, n_items, world)) for r in range(world)]
    order = [i for r in range(world) for i in range(r, n_items, world)]
    return counts, order


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> List[int]:
    """Long-form chunks are dealt round-robin: chunk w belongs to rank w % world."""
    return [w for w in range(n_chunks) if w % world == rank]


def all_gather_chunk_results(mine: Dict[int, Tuple[torch.Tensor, torch.Tensor]], n_chunks: int, m_chunk: int,
                             chunk_lens: Sequence[int]) -> Dict[int, Tuple[torch.Tensor, torch.Tensor]]:
    """Long-form path on several ranks: every rank clustered the chunks of `chunks_of_rank`; `mine[w]` = (labels int32 [m_w],
    within-cluster affinity mass float32 [m_w]) of chunk w on this rank's device.  One all_gather_into_tensor of a
    [chunks per rank, 2, m_chunk] int32 tensor (the mass bit-cast) gives every rank every chunk -- 8 bytes per window, no
    pickling through the host (round 1 used all_gather_object on the merged embeddings and index lists)."""
    rank, world = rank_world()
    per_rank = -(-n_chunks // world)
    ref = next(iter(mine.values()))[0] if mine else None
    device = ref.device if ref is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    buf = torch.zeros(per_rank, 2, m_chunk, dtype=torch.int32, device=device)
    for slot, w in enumerate(chunks_of_rank(n_chunks, rank, world)):
        y32, mass = mine[w]
        buf[slot, 0, : y32.numel()] = y32.to(torch.int32)
        buf[slot, 1, : mass.numel()] = mass.contiguous().view(torch.int32)
    out = torch.empty(world * per_rank, 2, m_chunk, dtype=torch.int32, device=device)
    torch.distributed.all_gather_into_tensor(out, buf)
    result = {}
    for w in range(n_chunks):
        row = out[(w % world) * per_rank + w // world]
        result[w] = (row[0, : chunk_lens[w]], row[1, : chunk_lens[w]].view(torch.float32))
    return result
