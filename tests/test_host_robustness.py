"""Property tests (hypothesis, CPU only) of the host logic around the GPU path: the RIFF parser never fails in an uncontrolled way,
the long-audio packing covers every segment once and its merge restores time order, the VAD end-point detector always returns
ordered, non-overlapping segments inside the recording, and the mini-sentence split of the punctuation model is a partition."""
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from funasr_b200 import _abi
from funasr_b200.audio import parse_wav_header
from funasr_b200.long_audio import merge_results, pack_segments
from funasr_b200.punc import split_to_mini_sentence, split_words
from funasr_b200.vad import chunk_frame_counts, detect_segments, merge_vad, num_frames


def _wav(tag, bits, ch, rate, payload, extra=b""):
    blk = ch * bits // 8
    fmt = struct.pack("<HHIIHH", tag, ch, rate, rate * blk, blk, bits)
    return b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + len(extra) + 8 + len(payload)) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra + \
        b"data" + struct.pack("<I", len(payload)) + payload


@settings(max_examples=300, deadline=None)
@given(st.binary(min_size=0, max_size=120))
def test_wav_parser_rejects_garbage_cleanly(blob):
    for data in (blob, b"RIFF" + blob, b"RIFF\x00\x00\x00\x00WAVE" + blob, b"RIFF\x00\x00\x00\x00WAVEfmt " + blob):
        try:
            code, ch, rate, off, nbytes = parse_wav_header(data)
        except (_abi.FunasrB200Error, struct.error):
            continue
        assert code in (0, 1, 2, 3, 4) and 0 <= off <= len(data) and 0 <= nbytes <= len(data) - off


@given(st.sampled_from([(1, 8), (1, 16), (1, 24), (1, 32), (3, 32)]), st.integers(1, 4), st.sampled_from([8000, 16000, 44100]),
       st.integers(0, 64), st.booleans())
def test_wav_parser_finds_the_data_chunk(fmt, ch, rate, frames, with_list_chunk):
    tag, bits = fmt
    payload = bytes(range(256)) * 4
    payload = payload[: frames * ch * bits // 8]
    extra = (b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00") if with_list_chunk else b""       # odd-sized chunk is padded to even
    code, c, r, off, nbytes = parse_wav_header(_wav(tag, bits, ch, rate, payload, extra))
    assert (c, r, nbytes) == (ch, rate, len(payload))
    assert code == {(1, 8): 4, (1, 16): 1, (1, 24): 2, (1, 32): 3, (3, 32): 0}[fmt]
    assert _wav(tag, bits, ch, rate, payload, extra)[off:off + nbytes] == payload


_segs = st.lists(st.tuples(st.integers(0, 600000), st.integers(1, 70000)), min_size=0, max_size=60).map(
    lambda xs: [[b, b + d] for b, d in sorted(xs)])


@settings(max_examples=200, deadline=None)
@given(_segs, st.integers(1, 400), st.integers(1, 80))
def test_pack_segments_partitions_the_sorted_order(segs, batch_size_s, threshold_s):
    order, packs = pack_segments(segs, batch_size_s, threshold_s)
    assert sorted(order) == list(range(len(segs)))
    dur = [segs[i][1] - segs[i][0] for i in order]
    assert dur == sorted(dur)                                              # shortest first (auto_model.py:918)
    covered = [i for b, e in packs for i in range(b, e)]
    assert covered == list(range(len(segs)))                               # contiguous, disjoint, complete
    for b, e in packs:
        if e - b > 1:                                                      # a pack of several segments respects the budget when it was closed
            assert max(dur[b:e - 1]) * (e - 1 - b) < max(batch_size_s * 1000, dur[0])
            assert all(d < threshold_s * 1000 for d in dur[b:e - 1])


@given(_segs)
def test_merge_results_restores_time_order_and_offsets(segs):
    per = [{"key": "k", "text": "t%d" % j, "timestamp": [[0, 10], [10, s[1] - s[0]]], "token_int": [j]} for j, s in enumerate(segs)]
    out = merge_results(per, segs)
    if not segs:
        assert out == {}
        return
    assert out["key"] == "k" and out["text"].split(" ") == ["t%d" % j for j in range(len(segs))]
    assert out["token_int"] == list(range(len(segs)))
    stamps = out["timestamp"]
    assert len(stamps) == 2 * len(segs)
    for j, s in enumerate(segs):
        assert stamps[2 * j] == [s[0], s[0] + 10] and stamps[2 * j + 1] == [s[0] + 10, s[1]]


@settings(max_examples=60, deadline=None)
@given(st.integers(720, 16000 * 150), st.integers(0, 2 ** 31 - 1), st.sampled_from([0.05, 0.5, 0.95]))   # >= 3 frames: below that the online frontend scores nothing
def test_vad_detector_segments_are_ordered_and_inside_the_recording(n_samples, seed, speech_share):
    g = np.random.default_rng(seed)
    t = num_frames(n_samples)
    # piecewise-constant speech / silence with random run lengths: silence posterior near 1 or near 0, energy high in speech
    sil, db, pos = np.empty(t, np.float32), np.empty(t, np.float32), 0
    while pos < t:
        run = int(g.integers(1, 400))
        speech = g.random() < speech_share
        sil[pos:pos + run] = 0.02 if speech else 0.98
        db[pos:pos + run] = 60.0 if speech else -20.0
        pos += run
    segs = detect_segments(sil.tolist(), db.tolist(), n_samples)
    dur_ms = n_samples // 16
    last_end = -1
    for s in segs:
        assert len(s) == 2 and 0 <= s[0] <= s[1] <= dur_ms + 60, (s, dur_ms)
        assert s[0] >= last_end
        last_end = s[1]
    assert sum(chunk_frame_counts(n_samples)) == t                         # every frame is delivered to the detector exactly once
    merged = merge_vad([list(s) for s in segs], 15000)
    assert sum(e - b for b, e in merged) >= sum(e - b for b, e in segs) or not segs


@given(st.text(alphabet=st.sampled_from(list("ab Z9你好世界，。 \t")), max_size=60), st.integers(2, 25))
def test_punctuation_word_split_and_mini_sentences_partition_the_text(text, limit):
    words = split_words(text)
    assert "".join(words) == "".join(text.split())                         # nothing lost, nothing invented, whitespace dropped
    for w in words:
        assert w and (all(len(c.encode()) == 1 for c in w) or len(w) == 1)  # ASCII runs or single wide characters
    if words:
        minis = split_to_mini_sentence(words, limit)
        assert [w for m in minis for w in m] == words and all(1 <= len(m) <= limit for m in minis)
