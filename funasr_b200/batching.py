"""Length-sorted bucketing of ragged utterance lists (BASELINE config 3: 512 synthetic 5–30 s utterances, bucketed
padding, utterance-sharded over the GPUs).

The reference batches long-audio segments the same way: sort by duration, then greedily pack while
`max_len x count` stays under a budget (funasr/auto/auto_model.py:918, :942-954).  Here the budget is a frame count
(`max_frames` padded LFR frames per batch) plus a cap on the number of utterances; batches are contiguous runs of the
sorted list so padding waste is minimal.  Pure host logic (no GPU) so it is unit-tested on CPU.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

from .engine import num_lfr_frames


def bucket_by_length(n_samples: Sequence[int], max_batch: int = 64, max_frames: int = 64 * 500) -> List[List[int]]:
    """Partition utterance indices into batches: sorted by length (longest first), each batch a contiguous run with at
    most `max_batch` utterances and `len(batch) * frames(longest) <= max_frames`.  Every index appears exactly once."""
    order = sorted(range(len(n_samples)), key=lambda i: (-int(n_samples[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_t = 0
    for i in order:
        t = max(1, num_lfr_frames(int(n_samples[i])))
        t_max = max(cur_t, t)
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * t_max > max_frames):
            batches.append(cur)
            cur, t_max = [], t
        cur.append(i)
        cur_t = t_max
    if cur:
        batches.append(cur)
    return batches


def padding_efficiency(n_samples: Sequence[int], batches: List[List[int]]) -> float:
    """Useful frames / padded frames over all batches (1.0 = no padding)."""
    useful = padded = 0
    for b in batches:
        ts = [num_lfr_frames(int(n_samples[i])) for i in b]
        useful += sum(ts)
        padded += len(ts) * max(ts)
    return useful / max(padded, 1)


def run_bucketed(wavs: Sequence, infer_batch: Callable[[List], List[List[int]]], max_batch: int = 64,
                 max_frames: int = 64 * 500) -> List[List[int]]:
    """Run `infer_batch(list_of_wavs) -> list_of_id_lists` over length buckets and return results in input order."""
    lens = [int(w.shape[-1]) for w in wavs]
    out: List[List[int]] = [[] for _ in wavs]
    for b in bucket_by_length(lens, max_batch, max_frames):
        res = infer_batch([wavs[i] for i in b])
        for i, r in zip(b, res):
            out[i] = r
    return out
