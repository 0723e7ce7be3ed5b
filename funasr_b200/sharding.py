"""Utterance sharding across the GPUs of one box and the single end-of-job all-gather of token ids.

Utterances are independent, so the data path has no collective: each rank (one process per
GPU, torchrun) decodes its shard; one ``all_gather`` of fixed-width int32 rows (ids padded to ``width``, plus the
lengths) over NCCL/NVLink returns every rank's result to all ranks.  The reference has no counterpart — its
inference is single-process (funasr/auto/auto_model.py:551-561).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_utterances(durations: Sequence[float], world_size: int) -> List[List[int]]:
    """Deal utterance indices to ranks so every rank gets near-equal work.

    Work per utterance grows a little faster than linearly in its duration (attention is O(T^2)), so indices are
    sorted by duration (longest first, the reference's own length-sorted batching precedent,
    auto_model.py:918) and dealt in snake order.  Deterministic; every index appears exactly once."""
    order = sorted(range(len(durations)), key=lambda i: (-float(durations[i]), i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for pos, idx in enumerate(order):
        rnd, r = divmod(pos, world_size)
        shards[r if rnd % 2 == 0 else world_size - 1 - r].append(idx)
    return shards


def gather_token_ids(local_ids: List[List[int]], local_index: Sequence[int], n_total: int, width: int = 512,
                     device=None, group=None) -> List[List[int]]:
    """All-gather every rank's greedy ids and return them in the original utterance order on every rank.

    local_ids[j] are the ids of utterance local_index[j].  Rows are padded to ``width`` (-1) so no size exchange
    round is needed; shards are padded to the largest shard with index -1 rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = (n_total + world - 1) // world
    # snake dealing gives shard sizes that differ by at most one; pad to `per`
    buf = torch.full((per, width + 2), -1, dtype=torch.int32)
    for j, (ids, idx) in enumerate(zip(local_ids, local_index)):
        n = min(len(ids), width)
        buf[j, 0] = idx
        buf[j, 1] = n
        if n:
            buf[j, 2:2 + n] = torch.tensor(ids[:n], dtype=torch.int32)
    if world == 1:
        gathered = buf[None]
    else:
        if device is not None:
            buf = buf.to(device)
        out = torch.empty((world * per, width + 2), dtype=torch.int32, device=buf.device)
        dist.all_gather_into_tensor(out, buf, group=group)     # the one collective of the job
        gathered = out.cpu().reshape(world, per, width + 2)
    result: List[List[int]] = [[] for _ in range(n_total)]
    for r in range(gathered.shape[0]):
        rows = gathered[r]
        for j in range(rows.shape[0]):
            idx = int(rows[j, 0])
            if idx >= 0:
                result[idx] = rows[j, 2:2 + int(rows[j, 1])].tolist()
    return result
