"""CIF-based token timestamps for the offline Paraformer path (host side, over the predictor kernel's outputs).

Restates what `Paraformer.inference(pred_timestamp=True)` does after the hot path (funasr/models/paraformer/model.py:673-683
calling funasr/utils/timestamp_tools.py:37-123 `ts_prediction_lfr6_standard`, which re-integrates the weights with
`cif_wo_hidden` :14-34): the CIF weights/fires the predictor kernel already produced (`fa_cif_predictor_forward`: alphas,
peaks) are turned into [start_ms, end_ms] per token.  Integer-millisecond results are bit-exact against the reference
(tests/test_timestamps.py pins them to golden vectors made by running the reference's own function).

Note on the call convention: model.py:674-676 passes `pre_peak_index[i]` as the function's first ("us_alphas") and
`alphas[i]` as its second ("us_peaks") argument; `paraformer_timestamps` reproduces exactly that call so results match the
reference's output, quirk included.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

EDGE_SILENCE_FRAMES = 5        # leading / trailing gap (in CIF frames) above which a <sil> segment is emitted
MAX_TOKEN_FRAMES = 12          # a token longer than this is cut and the remainder becomes <sil>


def cif_wo_hidden(alphas: np.ndarray, threshold: float) -> np.ndarray:
    """Integrate-and-fire trace of one utterance (timestamp_tools.py:14-34): fp32 running sum of the weights, reduced by
    `threshold` right after every frame where it reaches it; the returned trace holds the value BEFORE the reduction.
    Runs in the library's host code (fa_cif_wo_hidden_host: the same fp32 adds in the same order; the Python loop below,
    `cif_wo_hidden_py`, costs 1 ms per 1500 frames and BiCif / SeACo re-integrate every utterance)."""
    from . import _abi
    w = np.ascontiguousarray(np.asarray(alphas, dtype=np.float32).reshape(-1))
    trace = np.empty_like(w)
    _abi.check(_abi.load().fa_cif_wo_hidden_host(w.ctypes.data, w.size, float(np.float32(threshold)), trace.ctypes.data), "fa_cif_wo_hidden_host")
    return trace


def cif_wo_hidden_py(alphas: np.ndarray, threshold: float) -> np.ndarray:
    """The same trace as a plain numpy-scalar loop: the step-by-step specification `cif_wo_hidden` is tested against."""
    w = np.asarray(alphas, dtype=np.float32)
    trace = np.empty_like(w)
    level = np.float32(0.0)
    thr = np.float32(threshold)
    for t, a in enumerate(w):
        level = np.float32(level + a)
        trace[t] = level
        if level >= thr:
            level = np.float32(level - np.float32(1.0) * thr)
    return trace


def _fire_positions(trace: np.ndarray, shift: float) -> np.ndarray:
    return np.flatnonzero(trace >= np.float32(1.0 - 1e-4)).astype(np.float64) + shift


def ts_prediction_lfr6_standard(us_alphas, us_peaks, char_list: Sequence[str], vad_offset: float = 0.0, force_time_shift: float = -1.5,
                                sil_in_str: bool = True, upsample_rate: int = 3, want_text: bool = True) -> Tuple[str, List[List[int]]]:
    """Same contract as timestamp_tools.py:37-123 for one utterance: (formatted string, [[start_ms, end_ms] per token]).
    Inputs are not modified (the reference renormalises its first argument in place, :68).  want_text=False skips the formatted
    string (the model classes drop it, model.py:402-407 / :674-680; formatting is half of this routine's host time at 64 utterances
    per 40 ms GPU step) and returns "" in its place."""
    tokens = list(char_list)
    if not tokens:
        return "", []
    if tokens[-1] == "</s>":
        tokens.pop()
    sec_per_frame = 10.0 * 6 / 1000 / upsample_rate
    first = np.array(us_alphas, dtype=np.float32)
    second = np.array(us_peaks, dtype=np.float32)
    weights = first.reshape(-1) if first.ndim == 1 else first[0]
    trace = second.reshape(-1) if second.ndim == 1 else second[0]
    fires = _fire_positions(trace, force_time_shift)
    if fires.size != len(tokens) + 1:
        # the fire count disagrees with the token count: rescale the weights to sum to tokens + 1 and re-integrate (:67-72)
        with np.errstate(invalid="ignore", divide="ignore"):          # all-zero weights: NaN like the reference, no fires below
            weights = weights / np.float32(weights.sum(dtype=np.float32) / np.float32(len(tokens) + 1))
        trace = cif_wo_hidden(weights, 1.0 - 1e-4)
        fires = _fire_positions(trace, force_time_shift)
    if fires.size == 0:
        return "", []                      # the reference raises IndexError here (:83); nothing to time-stamp
    n_frames = trace.shape[0]
    if not want_text:
        return "", _stamps_only(fires, n_frames, tokens, sec_per_frame, vad_offset)
    labels: List[str] = []
    spans: List[List[float]] = []          # seconds

    def emit(label, lo, hi):
        labels.append(label)
        spans.append([lo * sec_per_frame, hi * sec_per_frame])

    if fires[0] > EDGE_SILENCE_FRAMES:
        emit("<sil>", 0.0, fires[0])
    for i, (lo, hi) in enumerate(zip(fires[:-1], fires[1:])):
        label = tokens[i] if i < len(tokens) else ""
        if MAX_TOKEN_FRAMES >= 0 and hi - lo > MAX_TOKEN_FRAMES:
            cut = lo + MAX_TOKEN_FRAMES
            emit(label, lo, cut)
            emit("<sil>", cut, hi)
        else:
            emit(label, lo, hi)
    if n_frames - fires[-1] > EDGE_SILENCE_FRAMES:
        mid = (n_frames + fires[-1]) * 0.5
        if spans:
            spans[-1][1] = mid * sec_per_frame
        emit("<sil>", mid, n_frames)
    elif spans:
        spans[-1][1] = n_frames * sec_per_frame
    if vad_offset:
        shift_s = vad_offset / 1000.0
        spans = [[lo + shift_s, hi + shift_s] for lo, hi in spans]
    text = "" if not want_text else "".join("{} {} {};".format(lab, str(lo + 0.0005)[:5], str(hi + 0.0005)[:5])
                                            for lab, (lo, hi) in zip(labels, spans) if sil_in_str or lab != "<sil>")
    stamps = [[int(lo * 1000), int(hi * 1000)] for lab, (lo, hi) in zip(labels, spans) if lab != "<sil>"]
    return text, stamps


def _stamps_only(fires: np.ndarray, n_frames: int, tokens: Sequence[str], sec_per_frame: float, vad_offset: float) -> List[List[int]]:
    """The [[start_ms, end_ms]] list of `ts_prediction_lfr6_standard` without building labels / the string: the same double
    arithmetic element by element (start = lo * sec, end = (lo + 12 or hi) * sec, the trailing-edge rule on the LAST emitted span,
    + vad_offset / 1000, int(x * 1000)), vectorised over the tokens.  A token cut at MAX_TOKEN_FRAMES is followed by a <sil> span,
    which then is the last emitted span when it happens to the last token — the trailing-edge rule moves that <sil>, not the token.
    Spans beyond the token list (label "") are stamps too, a vocabulary entry spelled "<sil>" is dropped like an inserted one."""
    lo, hi = fires[:-1], fires[1:]
    n_span = lo.shape[0]
    if n_span == 0:
        return []                                               # only edge <sil> spans exist
    start = lo * sec_per_frame
    cut = (hi - lo > MAX_TOKEN_FRAMES) if MAX_TOKEN_FRAMES >= 0 else np.zeros(n_span, dtype=bool)
    end = np.where(cut, lo + MAX_TOKEN_FRAMES, hi) * sec_per_frame
    if not cut[-1]:                                             # the last emitted span is the last token itself
        if n_frames - fires[-1] > EDGE_SILENCE_FRAMES:
            end[-1] = ((n_frames + fires[-1]) * 0.5) * sec_per_frame
        else:
            end[-1] = n_frames * sec_per_frame
    if vad_offset:
        shift_s = vad_offset / 1000.0
        start, end = start + shift_s, end + shift_s
    out = np.stack([start * 1000, end * 1000], axis=1).astype(np.int64)          # int(): truncation toward zero, like astype
    if "<sil>" in tokens:
        keep = np.array([not (i < len(tokens) and tokens[i] == "<sil>") for i in range(n_span)], dtype=bool)
        out = out[keep]
    return out.tolist()


def paraformer_timestamps(peaks_row, alphas_row, tokens: Sequence[str], begin_time: float = 0.0, want_text: bool = True) -> Tuple[str, List[List[int]]]:
    """The exact call of paraformer/model.py:674-680: (pre_peak_index[i], alphas[i], tokens, vad_offset=begin_time,
    upsample_rate=1)."""
    return ts_prediction_lfr6_standard(peaks_row, alphas_row, list(tokens), vad_offset=begin_time, upsample_rate=1, want_text=want_text)
