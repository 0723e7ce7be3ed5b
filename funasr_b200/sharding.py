"""Utterance sharding across the GPUs of one box and the single end-of-job all-gather of token ids.

Utterances are independent, so the data path has no collective: each rank (one process per
GPU, torchrun) decodes its shard; one ``all_gather`` of fixed-width int32 rows (ids padded to ``width``, plus the
lengths) over NCCL/NVLink returns every rank's result to all ranks.  The reference has no counterpart — its
inference is single-process (funasr/auto/auto_model.py:551-561).

``ShardedRunner`` is the product entry point that composes the pieces: shard (duration-sorted snake deal) -> bucket
(length-sorted padded batches, batching.py) -> infer (the caller's batch function, e.g. the engine) -> gather (rows built on
the device, one collective, optionally asynchronous so that it overlaps the next job's compute).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .batching import bucket_by_length
from .engine import num_lfr_frames


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


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def rows_to_lists(gathered: torch.Tensor, n_total: int) -> List[List[int]]:
    """[world * per, width + 2] int32 rows (index, length, ids...) on the host -> id lists in original utterance order."""
    g = gathered.cpu()
    idx = g[:, 0].tolist()
    lens = g[:, 1].tolist()
    result: List[List[int]] = [[] for _ in range(n_total)]
    for j, (i, n) in enumerate(zip(idx, lens)):
        if i >= 0:
            result[i] = g[j, 2:2 + n].tolist()
    return result


def gather_token_ids(local_ids: List[List[int]], local_index: Sequence[int], n_total: int, width: Optional[int] = None,
                     device=None, group=None) -> List[List[int]]:
    """All-gather every rank's greedy ids (host lists) and return them in the original utterance order on every rank.

    local_ids[j] are the ids of utterance local_index[j].  Rows are padded to ``width`` (-1); shards are padded to the largest
    shard with index -1 rows.  ``width=None`` sizes the rows from the longest id list of ANY rank (one extra all_reduce(MAX) of
    a single integer); an explicit width that is too small raises instead of truncating — the single-GPU path returns every id,
    so the multi-GPU path must too."""
    world, _ = _world(group)
    per = (n_total + world - 1) // world
    longest = max((len(r) for r in local_ids), default=0)
    if width is None:
        if world > 1:
            t = torch.tensor([longest], dtype=torch.int32, device=device if device is not None else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            longest = int(t.item())
        width = max(longest, 1)
    elif longest > width:
        raise ValueError("gather_token_ids: an utterance has %d ids but rows are %d wide; pass width=None or a larger width" % (longest, width))
    flat = torch.full((per, width + 2), -1, dtype=torch.int32)
    if local_ids:
        flat[: len(local_ids), 0] = torch.tensor(list(local_index), dtype=torch.int32)
        flat[: len(local_ids), 1] = torch.tensor([len(r) for r in local_ids], dtype=torch.int32)
        pad = torch.nn.utils.rnn.pad_sequence([torch.tensor(r, dtype=torch.int32) for r in local_ids], batch_first=True, padding_value=-1)
        flat[: len(local_ids), 2:2 + pad.shape[1]] = pad
    if world == 1:
        return rows_to_lists(flat, n_total)
    buf = flat.to(device) if device is not None else flat
    out = torch.empty((world * per, width + 2), dtype=torch.int32, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)     # the one collective of the job
    return rows_to_lists(out, n_total)


class ShardedRunner:
    """shard -> bucket -> infer -> gather for one job (a list of utterances known to every rank).

    infer_batch(list_of_wavs) -> (ids [b, n] int32 on the device, padded with -1; lens [b] int32 on the device): the greedy ids
    of one padded batch (e.g. ``lambda ws: engine_outputs(ws)["ids_dev"], ["ids_lens_dev"]``).  Result rows are assembled ON THE
    DEVICE (index, length, ids; three strided copies per bucket, no Python loop over tokens) and exchanged with one
    ``all_gather_into_tensor`` per job.  ``gather_async`` issues the collective on NCCL's own stream so the next job's kernels
    are not ordered behind it; ``finish`` waits and de-permutes."""

    def __init__(self, infer_batch: Callable[[List[torch.Tensor]], Tuple[torch.Tensor, torch.Tensor]], device, max_batch: int = 64,
                 max_frames: int = 64 * 500, group=None, extra_ids: int = 1):
        """extra_ids: id columns beyond the utterance's LFR frame count T a result row may need — 1 for the CIF models (tokens
        <= T + 1, the tail token), 4 for SenseVoiceSmall (CTC ids over T + 4 frames: the four prepended query frames)."""
        self.infer_batch = infer_batch
        self.extra_ids = int(extra_ids)
        self.device = torch.device(device)
        self.max_batch, self.max_frames, self.group = max_batch, max_frames, group
        self.world, self.rank = _world(group)
        self._bufs = {}

    # ---- planning is pure host logic, identical on every rank (all ranks know every utterance's length)
    def plan(self, n_samples: Sequence[int]):
        shards = shard_utterances([float(n) for n in n_samples], self.world)
        mine = shards[self.rank]
        local_lens = [int(n_samples[i]) for i in mine]
        buckets = [[mine[j] for j in b] for b in bucket_by_length(local_lens, self.max_batch, self.max_frames)]
        per = max(len(s) for s in shards)
        width = max(num_lfr_frames(int(n)) for n in n_samples) + self.extra_ids if len(n_samples) else 1   # bound on ids per utterance
        return {"shards": shards, "mine": mine, "buckets": buckets, "per": per, "width": width, "n_total": len(n_samples)}

    def _rows_buffer(self, per: int, width: int, slot: int) -> torch.Tensor:
        key = (per, width, slot)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty((per, width + 2), dtype=torch.int32, device=self.device)
        return t

    def pack_rows(self, rows: torch.Tensor, at: int, index: Sequence[int], ids: torch.Tensor, lens: torch.Tensor) -> int:
        """rows[at : at+b] = (index, lens, ids) — device-side copies."""
        b, n = ids.shape
        if n > rows.shape[1] - 2:
            raise ValueError("ShardedRunner: %d id columns do not fit rows of width %d" % (n, rows.shape[1] - 2))
        view = rows[at: at + b]
        view[:, 0].copy_(torch.tensor(list(index), dtype=torch.int32), non_blocking=True)
        view[:, 1].copy_(lens)
        view[:, 2:2 + n].copy_(ids)
        return at + b

    def run_local(self, wavs: Sequence[torch.Tensor], plan: dict, slot: int = 0) -> torch.Tensor:
        rows = self._rows_buffer(plan["per"], plan["width"], slot)
        rows.fill_(-1)
        at = 0
        for b in plan["buckets"]:
            ids, lens = self.infer_batch([wavs[i] for i in b])
            at = self.pack_rows(rows, at, b, ids, lens)
        return rows

    def gather_async(self, rows: torch.Tensor, slot: int = 0):
        if self.world == 1:
            return (rows, None)
        key = ("out", rows.shape[0], rows.shape[1], slot)
        out = self._bufs.get(key)
        if out is None:
            out = self._bufs[key] = torch.empty((self.world * rows.shape[0], rows.shape[1]), dtype=torch.int32, device=self.device)
        work = dist.all_gather_into_tensor(out, rows, group=self.group, async_op=True)     # the one collective of the job
        return (out, work)

    def finish(self, handle, n_total: int) -> List[List[int]]:
        out, work = handle
        if work is not None:
            work.wait()
        return rows_to_lists(out, n_total)

    def run(self, wavs: Sequence[torch.Tensor]) -> List[List[int]]:
        """The whole job: every rank passes the same list; every rank returns all id lists in input order."""
        plan = self.plan([int(w.shape[-1]) for w in wavs])
        rows = self.run_local(wavs, plan)
        return self.finish(self.gather_async(rows), plan["n_total"])
