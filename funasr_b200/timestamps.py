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

START_END_THRESHOLD = 5
MAX_TOKEN_DURATION = 12


def cif_wo_hidden(alphas: np.ndarray, threshold: float) -> np.ndarray:
    """timestamp_tools.py:14-34 for one utterance: running fp32 integral, minus `threshold` after every fire; returns the
    integral BEFORE the reset at every frame."""
    a = np.asarray(alphas, dtype=np.float32)
    fires = np.empty_like(a)
    integrate = np.float32(0.0)
    thr = np.float32(threshold)
    for t in range(a.shape[0]):
        integrate = np.float32(integrate + a[t])
        fires[t] = integrate
        if integrate >= thr:
            integrate = np.float32(integrate - np.float32(1.0) * thr)
    return fires


def ts_prediction_lfr6_standard(us_alphas, us_peaks, char_list: Sequence[str], vad_offset: float = 0.0, force_time_shift: float = -1.5,
                                sil_in_str: bool = True, upsample_rate: int = 3) -> Tuple[str, List[List[int]]]:
    """Same contract as timestamp_tools.py:37-123 (one utterance).  Does not modify its inputs (the reference renormalises
    its `us_alphas` argument in place, :68)."""
    char_list = list(char_list)
    if not len(char_list):
        return "", []
    time_rate = 10.0 * 6 / 1000 / upsample_rate
    alphas = np.array(us_alphas, dtype=np.float32).reshape(-1) if np.ndim(us_alphas) == 1 else np.array(us_alphas, dtype=np.float32)[0]
    peaks = np.array(us_peaks, dtype=np.float32).reshape(-1) if np.ndim(us_peaks) == 1 else np.array(us_peaks, dtype=np.float32)[0]
    if char_list[-1] == "</s>":
        char_list = char_list[:-1]
    thr = np.float32(1.0 - 1e-4)
    fire_place = np.nonzero(peaks >= thr)[0].astype(np.float64) + force_time_shift
    if len(fire_place) != len(char_list) + 1:
        alphas = alphas / np.float32(alphas.sum(dtype=np.float32) / np.float32(len(char_list) + 1))
        peaks = cif_wo_hidden(alphas, 1.0 - 1e-4)
        fire_place = np.nonzero(peaks >= thr)[0].astype(np.float64) + force_time_shift
    if len(fire_place) == 0:
        return "", []                      # the reference raises IndexError here (:83); nothing to time-stamp
    num_frames = peaks.shape[0]
    stamps: List[List[float]] = []
    chars: List[str] = []
    if fire_place[0] > START_END_THRESHOLD:
        stamps.append([0.0, fire_place[0] * time_rate])
        chars.append("<sil>")
    for i in range(len(fire_place) - 1):
        chars.append(char_list[i] if i < len(char_list) else "")
        if MAX_TOKEN_DURATION < 0 or fire_place[i + 1] - fire_place[i] <= MAX_TOKEN_DURATION:
            stamps.append([fire_place[i] * time_rate, fire_place[i + 1] * time_rate])
        else:
            split = fire_place[i] + MAX_TOKEN_DURATION
            stamps.append([fire_place[i] * time_rate, split * time_rate])
            stamps.append([split * time_rate, fire_place[i + 1] * time_rate])
            chars.append("<sil>")
    if num_frames - fire_place[-1] > START_END_THRESHOLD:
        end = (num_frames + fire_place[-1]) * 0.5
        if stamps:
            stamps[-1][1] = end * time_rate
        stamps.append([end * time_rate, num_frames * time_rate])
        chars.append("<sil>")
    elif stamps:
        stamps[-1][1] = num_frames * time_rate
    if vad_offset:
        for s in stamps:
            s[0] += vad_offset / 1000.0
            s[1] += vad_offset / 1000.0
    txt = ""
    for ch, s in zip(chars, stamps):
        if not sil_in_str and ch == "<sil>":
            continue
        txt += "{} {} {};".format(ch, str(s[0] + 0.0005)[:5], str(s[1] + 0.0005)[:5])
    res = [[int(s[0] * 1000), int(s[1] * 1000)] for ch, s in zip(chars, stamps) if ch != "<sil>"]
    return txt, res


def paraformer_timestamps(peaks_row, alphas_row, tokens: Sequence[str], begin_time: float = 0.0) -> Tuple[str, List[List[int]]]:
    """The exact call of paraformer/model.py:674-680: (pre_peak_index[i], alphas[i], tokens, vad_offset=begin_time,
    upsample_rate=1)."""
    return ts_prediction_lfr6_standard(peaks_row, alphas_row, list(tokens), vad_offset=begin_time, upsample_rate=1)
