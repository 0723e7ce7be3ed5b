"""FSMN-VAD for the long-audio path: the end-point detector over per-frame silence posteriors (host side) and the frame
bookkeeping of the reference's chunked, stateful frontend.

The reference (funasr/models/fsmn_vad_streaming/model.py) evaluates its FSMN on 60 s chunks of the waveform through
WavFrontendOnline (stateful framing + LFR 5/1, frontends/wav_frontend.py:259-660) and then walks over the frames ONE AT A TIME
in Python (GetFrameState :761-823, WindowDetector :218-320, DetectOneFrame :1158-1302, the On* callbacks :641-736) to turn the
silence posterior of each 10 ms frame into [start_ms, end_ms] segments.  Here the per-frame arithmetic (Fbank + LFR + CMVN, the
FSMN stack, softmax, frame energies) runs on the GPU over the WHOLE waveform in one pass (csrc/vad.cu: the FSMN is causal, so
chunked evaluation with a cache equals one pass) and this module restates the sequential decision logic — integer / threshold
logic that the reference also runs on the host — including the parts that depend on the 60 s chunking: which frames arrive with
which chunk (`chunk_frame_counts`) and the per-chunk dynamic end-silence schedule (model.py:1013-1067, dynamic_vad.py:37-44).

Pinned by tests/test_vad_host.py against golden segments produced by the unmodified reference from the same per-frame scores.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# VadStateMachine / FrameState / AudioChangeState values of the reference (model.py:44-63)
_START_NOT_DETECTED, _IN_SPEECH, _END_DETECTED = 1, 2, 3
_SPEECH, _SIL = 1, 0
_SP2SP, _SP2SIL, _SIL2SIL, _SIL2SP = 0, 1, 2, 3

# (accumulated speech ms limit, end-silence ms) — model.py:33-41 (the schedule FsmnVADStreaming.inference uses; dynamic_vad.py has its own)
DEFAULT_SILENCE_SCHEDULE = [(10000, 2000), (20000, 1000), (30000, 800), (40000, 600), (50000, 400), (60000, 200), (float("inf"), 100)]


@dataclass
class VadOptions:
    """VADXOptions (model.py:71-174) with the template.yaml values of the released fsmn-vad model."""
    sample_rate: int = 16000
    detect_mode: int = 1
    max_end_silence_time: int = 800
    max_start_silence_time: int = 3000
    window_size_ms: int = 200
    sil_to_speech_time_thres: int = 150
    speech_to_sil_time_thres: int = 150
    speech_2_noise_ratio: float = 1.0
    do_extend: int = 1
    lookback_time_start_point: int = 200
    lookahead_time_end_point: int = 100
    max_single_segment_time: int = 60000
    snr_thres: float = -100.0
    noise_frame_num_used_for_snr: int = 100
    decibel_thres: float = -100.0
    speech_noise_thres: float = 0.6
    fe_prior_thres: float = 1e-4
    sil_pdf_ids: List[int] = field(default_factory=lambda: [0])
    frame_in_ms: int = 10
    frame_length_ms: int = 25

    @classmethod
    def from_conf(cls, conf: dict) -> "VadOptions":
        names = set(cls.__dataclass_fields__)
        return cls(**{k: v for k, v in (conf or {}).items() if k in names})


def num_frames(n_samples: int, frame_len: int = 400, shift: int = 160) -> int:
    return (n_samples - frame_len) // shift + 1 if n_samples >= frame_len else 0


def chunk_frame_counts(n_samples: int, chunk_ms: int = 60000, fs: int = 16000, lfr_m: int = 5, frame_len: int = 400, shift: int = 160) -> List[int]:
    """Score frames the reference's pipeline hands to the detector per waveform chunk (the last entry is the final chunk).

    FsmnVADStreaming.inference (model.py:1003-1064) cuts the waveform into `len // stride + 1` chunks (the last one may be empty);
    WavFrontendOnline.forward_fbank (wav_frontend.py:385-447) frames `leftover + chunk` and keeps the unframed tail;
    apply_lfr (:345-374, lfr_n = 1) starts from (lfr_m-1)/2 copies of the first frame, withholds lfr_m-1 frames of right
    context on non-final calls and pads with copies of the last frame on the final call.  Only integer arithmetic here."""
    stride = int(chunk_ms * fs / 1000)
    n_chunks = n_samples // stride + 1
    half = (lfr_m - 1) // 2
    out, leftover, cached = [], 0, None            # cached: frames held in lfr_splice_cache (None before the first frame)
    for i in range(n_chunks):
        final = i == n_chunks - 1
        new = min(stride, max(0, n_samples - i * stride))
        total = leftover + new
        f = int((total - frame_len) / shift + 1) if total >= frame_len else 0
        f = f if f >= 1 else 0
        leftover = total - f * shift
        emitted = 0
        if f > 0:
            if cached is None:
                cached = half
            t = cached + f
            if t >= lfr_m:
                emitted = t - half if final else t - (lfr_m - 1)          # T_lfr on the final call, last_idx otherwise
                cached = t - emitted if not final else t - min(t - 1, emitted)
            else:
                cached = t                                                  # too few frames: everything stays cached, nothing is scored
        elif final and cached:
            emitted = max(int(math.ceil(cached - half)), 0)                 # flush of the cached frames (wav_frontend.py:591-603)
        out.append(int(emitted))
    return out


class _Window:
    """WindowDetector (model.py:218-320): a sliding count of speech frames with hysteresis."""

    def __init__(self, o: VadOptions):
        self.n = int(o.window_size_ms / o.frame_in_ms)
        self.s2s = int(o.sil_to_speech_time_thres / o.frame_in_ms)
        self.sp2sil = int(o.speech_to_sil_time_thres / o.frame_in_ms)
        self.reset()

    def reset(self):
        self.pos, self.total, self.buf, self.pre = 0, 0, [0] * self.n, _SIL

    def step(self, state: int) -> int:
        self.total += state - self.buf[self.pos]
        self.buf[self.pos] = state
        self.pos = (self.pos + 1) % self.n
        if self.pre == _SIL and self.total >= self.s2s:
            self.pre = _SPEECH
            return _SIL2SP
        if self.pre == _SPEECH and self.total <= self.sp2sil:
            self.pre = _SIL
            return _SP2SIL
        return _SIL2SIL if self.pre == _SIL else _SP2SP


class VadDetector:
    """The stateful end-point detector (Stats + the Detect* / On* methods of FsmnVADStreaming).  Feed it the silence
    posterior and the energy of consecutive frames with `process_block`; segments are collected in `segments` ([start_ms,
    end_ms], complete ones; `take_new()` returns those not yet reported, like forward() :875-905 in offline mode)."""

    def __init__(self, opts: Optional[VadOptions] = None, speech_noise_thres: Optional[float] = None):
        self.o = opts or VadOptions()
        self.win = _Window(self.o)
        o = self.o
        self.frm_cnt = 0
        self.data_buf_start_frame = 0
        self.latest_confirmed_speech_frame = 0
        self.lastest_confirmed_silence_frame = -1
        self.continous_silence_frame_count = 0
        self.state = _START_NOT_DETECTED
        self.confirmed_start_frame = -1
        self.confirmed_end_frame = -1
        self.number_end_time_detected = 0
        self.noise_average_decibel = -100.0
        self.max_end_sil_frame_cnt_thresh = o.max_end_silence_time - o.speech_to_sil_time_thres
        self.speech_noise_thres = o.speech_noise_thres if speech_noise_thres is None else speech_noise_thres
        self.out: List[List] = []                  # [start_ms, end_ms, has_start, has_end]
        self.out_offset = 0
        self.latency = self.win.n + (int(o.lookback_time_start_point / o.frame_in_ms) if o.do_extend else 0)   # LatencyFrmNumAtStartPoint
        self._sil: List[float] = []
        self._db: List[float] = []

    # ---- GetFrameState (model.py:761-823), arithmetic in Python floats exactly as written there
    def _frame_state(self, t: int) -> int:
        o = self.o
        if t >= len(self._db):
            return _SIL
        cur_decibel = self._db[t]
        cur_snr = cur_decibel - self.noise_average_decibel
        if cur_decibel < o.decibel_thres:
            return _SIL
        sum_score = self._sil[t]
        noise_prob = math.log(sum_score) * o.speech_2_noise_ratio
        sum_score = 1.0 - sum_score
        speech_prob = math.log(sum_score)
        if math.exp(speech_prob) >= math.exp(noise_prob) + self.speech_noise_thres:
            return _SPEECH if (cur_snr >= o.snr_thres and cur_decibel >= o.decibel_thres) else _SIL
        if self.noise_average_decibel < -99.9:
            self.noise_average_decibel = cur_decibel
        else:
            self.noise_average_decibel = (cur_decibel + self.noise_average_decibel * (o.noise_frame_num_used_for_snr - 1)) / o.noise_frame_num_used_for_snr
        return _SIL

    # ---- buffer bookkeeping reduced to what decides integer outcomes (PopDataBufTillFrame / PopDataToOutputBuf :552-639):
    #      the sample buffer always holds the frames received so far, so popping "till frame" only advances the start frame
    def _pop_till(self, frame_idx: int):
        if self.data_buf_start_frame < frame_idx:
            self.data_buf_start_frame = frame_idx

    def _pop_to_output(self, start_frm: int, frm_cnt: int, first_is_start: bool, last_is_end: bool):
        ms = self.o.frame_in_ms
        self._pop_till(start_frm)
        if not self.out or first_is_start:
            self.out.append([start_frm * ms, start_frm * ms, False, False])
        seg = self.out[-1]
        self.data_buf_start_frame += frm_cnt
        seg[1] = (start_frm + frm_cnt) * ms
        if first_is_start:
            seg[2] = True
        if last_is_end:
            seg[3] = True

    def _on_silence(self, frame: int):
        self.lastest_confirmed_silence_frame = frame
        if self.state == _START_NOT_DETECTED:
            self._pop_till(frame)

    def _on_voice(self, frame: int):
        self.latest_confirmed_speech_frame = frame
        self._pop_to_output(frame, 1, False, False)

    def _on_voice_start(self, start_frame: int, fake: bool = False):
        if self.confirmed_start_frame == -1:
            self.confirmed_start_frame = start_frame
        if not fake and self.state == _START_NOT_DETECTED:
            self._pop_to_output(self.confirmed_start_frame, 1, True, False)

    def _on_voice_end(self, end_frame: int, fake: bool):
        for t in range(self.latest_confirmed_speech_frame + 1, end_frame):
            self._on_voice(t)
        if self.confirmed_end_frame == -1:
            self.confirmed_end_frame = end_frame
        if not fake:
            self._pop_to_output(self.confirmed_end_frame, 1, False, True)
        self.number_end_time_detected += 1

    def _reset_detection(self):
        self.continous_silence_frame_count = 0
        self.latest_confirmed_speech_frame = 0
        self.lastest_confirmed_silence_frame = -1
        self.confirmed_start_frame = -1
        self.confirmed_end_frame = -1
        self.state = _START_NOT_DETECTED
        self.win.reset()

    # ---- DetectOneFrame (model.py:1158-1302)
    def _detect(self, frame_state: int, cur: int, is_final: bool):
        o = self.o
        ms = o.frame_in_ms
        if frame_state == _SPEECH and not (math.fabs(1.0) > o.fe_prior_thres):
            frame_state = _SIL
        change = self.win.step(frame_state)
        too_long = lambda: cur - self.confirmed_start_frame + 1 > o.max_single_segment_time / ms

        def end_or_continue():
            if too_long():
                self._on_voice_end(cur, False)
                self.state = _END_DETECTED
            elif not is_final:
                self._on_voice(cur)
            else:
                self._on_voice_end(cur, False)                       # MaybeOnVoiceEndIfLastFrame
                self.state = _END_DETECTED

        if change == _SIL2SP:
            self.continous_silence_frame_count = 0
            if self.state == _START_NOT_DETECTED:
                start = max(self.data_buf_start_frame, cur - self.latency)
                self._on_voice_start(start)
                self.state = _IN_SPEECH
                for t in range(start + 1, cur + 1):
                    self._on_voice(t)
            elif self.state == _IN_SPEECH:
                for t in range(self.latest_confirmed_speech_frame + 1, cur):
                    self._on_voice(t)
                end_or_continue()
        elif change == _SP2SIL:
            self.continous_silence_frame_count = 0
            if self.state == _IN_SPEECH:
                end_or_continue()
        elif change == _SP2SP:
            self.continous_silence_frame_count = 0
            if self.state == _IN_SPEECH:
                end_or_continue()
        elif change == _SIL2SIL:
            self.continous_silence_frame_count += 1
            if self.state == _START_NOT_DETECTED:
                if (o.detect_mode == 0 and self.continous_silence_frame_count * ms > o.max_start_silence_time) or \
                        (is_final and self.number_end_time_detected == 0):
                    for t in range(self.lastest_confirmed_silence_frame + 1, cur):
                        self._on_silence(t)
                    self._on_voice_start(0, True)
                    self._on_voice_end(0, True)
                    self.state = _END_DETECTED
                elif cur >= self.latency:
                    self._on_silence(cur - self.latency)
            elif self.state == _IN_SPEECH:
                if self.continous_silence_frame_count * ms >= self.max_end_sil_frame_cnt_thresh:
                    lookback = int(self.max_end_sil_frame_cnt_thresh / ms)
                    if o.do_extend:
                        lookback -= int(o.lookahead_time_end_point / ms)
                        lookback -= 1
                        lookback = max(0, lookback)
                    self._on_voice_end(cur - lookback, False)
                    self.state = _END_DETECTED
                elif too_long():
                    self._on_voice_end(cur, False)
                    self.state = _END_DETECTED
                elif o.do_extend and not is_final:
                    if self.continous_silence_frame_count <= int(o.lookahead_time_end_point / ms):
                        self._on_voice(cur)
                elif is_final:
                    self._on_voice_end(cur, False)
                    self.state = _END_DETECTED
        if self.state == _END_DETECTED and o.detect_mode == 1:
            self._reset_detection()

    def process_block(self, sil_prob: Sequence[float], decibel: Sequence[float], is_final: bool):
        """One forward() of the reference (model.py:825-909) over the frames of one chunk: DetectCommonFrames /
        DetectLastFrames, i.e. every frame in order, the last one flagged final on the final chunk."""
        n = len(sil_prob)
        if n == 0:
            return
        self._sil.extend(float(x) for x in sil_prob)
        self._db.extend(float(x) for x in decibel)
        first = self.frm_cnt
        self.frm_cnt += n
        if self.state == _END_DETECTED:
            return
        for k in range(n):
            t = first + k
            self._detect(self._frame_state(t), t, is_final and k == n - 1)

    def take_new(self, is_final: bool) -> List[List[int]]:
        """Segments completed since the last call — forward() :875-905, is_streaming_input False."""
        got = []
        for i in range(self.out_offset, len(self.out)):
            s, e, has_s, has_e = self.out[i]
            if not is_final and (not has_s or not has_e):
                continue
            got.append([s, e])
            self.out_offset += 1
        return got


def detect_segments(sil_prob: Sequence[float], decibel: Sequence[float], n_samples: int, opts: Optional[VadOptions] = None,
                    chunk_ms: int = 60000, dynamic_silence: Optional[bool] = None, silence_schedule=None,
                    speech_noise_thres: Optional[float] = None, max_end_silence_time: Optional[int] = None) -> List[List[int]]:
    """FsmnVADStreaming.inference for one whole waveform (offline: is_final on the last chunk) given the per-frame silence
    posterior and energy of ALL frames -> [[start_ms, end_ms], ...].  Reproduces the reference's chunk-by-chunk delivery of
    frames and its per-chunk dynamic end-silence threshold (model.py:1003-1067)."""
    o = opts or VadOptions()
    if dynamic_silence is None:                                        # model.py:1006-1009: an explicit threshold switches the schedule off
        dynamic_silence = max_end_silence_time is None
    if max_end_silence_time is not None:
        o = VadOptions(**{**o.__dict__, "max_end_silence_time": max_end_silence_time})
    schedule = silence_schedule or DEFAULT_SILENCE_SCHEDULE
    det = VadDetector(o, speech_noise_thres)
    counts = chunk_frame_counts(n_samples, chunk_ms, o.sample_rate, 5, int(o.frame_length_ms * o.sample_rate / 1000),
                                int(o.frame_in_ms * o.sample_rate / 1000))
    segments: List[List[int]] = []
    accumulated_ms, in_speech, pos = 0, False, 0
    for i, cnt in enumerate(counts):
        final = i == len(counts) - 1
        if dynamic_silence:
            if det.state == _IN_SPEECH or in_speech:
                accumulated_ms += chunk_ms
                in_speech = True
            for limit_ms, silence_ms in schedule:
                if accumulated_ms <= limit_ms:
                    det.max_end_sil_frame_cnt_thresh = max(silence_ms - o.speech_to_sil_time_thres, 0)
                    det.speech_noise_thres = 0.5
                    break
        if cnt <= 0:
            continue                                                   # forward() returns [] for an empty feature block (:842-843)
        det.process_block(sil_prob[pos: pos + cnt], decibel[pos: pos + cnt], final)
        pos += cnt
        new = det.take_new(final)
        if new:
            segments.extend(new)
            if dynamic_silence:
                accumulated_ms, in_speech = 0, False
    return segments


def detect_segments_native(sil_prob, decibel, n_samples: int, opts: Optional[VadOptions] = None, chunk_ms: int = 60000,
                           dynamic_silence: Optional[bool] = None, silence_schedule=None, speech_noise_thres: Optional[float] = None,
                           max_end_silence_time: Optional[int] = None) -> List[List[int]]:
    """Same contract and same results as `detect_segments`, computed by the C++ state machine in the library
    (csrc/vad_detector.cpp: fa_vad_detect_segments) — ~200x faster than walking the frames in Python, which matters once the GPU
    scores an hour of audio in milliseconds.  sil_prob / decibel: anything numpy can view as a 1-D float array (fp32 values are
    widened to double exactly, like `float(x)` above)."""
    import ctypes as C

    import numpy as np

    from . import _abi
    o = opts or VadOptions()
    if dynamic_silence is None:
        dynamic_silence = max_end_silence_time is None
    if max_end_silence_time is not None:
        o = VadOptions(**{**o.__dict__, "max_end_silence_time": max_end_silence_time})
    sil = np.ascontiguousarray(np.asarray(sil_prob, dtype=np.float64).reshape(-1))
    db = np.ascontiguousarray(np.asarray(decibel, dtype=np.float64).reshape(-1))
    if sil.shape != db.shape:
        raise _abi.FunasrB200Error("detect_segments_native: sil_prob and decibel must hold one value per frame each")
    c = _abi.FaVadOptions()
    for name, ctype in _abi.FaVadOptions._fields_:
        v = getattr(o, name)
        if ctype is C.c_int32:
            if float(v) != int(v):                      # a fractional millisecond option: only the Python walk defines what that means
                return detect_segments(sil.tolist(), db.tolist(), n_samples, opts, chunk_ms, dynamic_silence, silence_schedule, speech_noise_thres,
                                       max_end_silence_time)
            v = int(v)
        setattr(c, name, v)
    sched = np.array([[-1.0 if math.isinf(lim) else float(lim), float(ms)] for lim, ms in (silence_schedule or DEFAULT_SILENCE_SCHEDULE)],
                     dtype=np.float64).reshape(-1)
    lib = _abi.load()
    cap = 64
    while True:
        out = np.empty((cap, 2), dtype=np.int32)
        n = int(lib.fa_vad_detect_segments(sil.ctypes.data, db.ctypes.data, sil.size, int(n_samples), C.byref(c), int(chunk_ms),
                                           1 if dynamic_silence else 0, sched.ctypes.data, sched.size // 2,
                                           float("nan") if speech_noise_thres is None else float(speech_noise_thres), out.ctypes.data, cap))
        if n < 0:
            raise _abi.FunasrB200Error("fa_vad_detect_segments failed (%d): posteriors must lie inside (0, 1)" % n)
        if n <= cap:
            return out[:n].tolist()
        cap = n


def merge_vad(vad_result: List[List[int]], max_length: int = 15000, min_length: int = 0) -> List[List[int]]:
    """funasr/utils/vad_utils.py:57-91: merge consecutive segments up to max_length ms (used by `merge_vad=True`)."""
    if len(vad_result) <= 1:
        return vad_result
    steps = sorted(set([t[0] for t in vad_result] + [t[1] for t in vad_result]))
    if not steps:
        return []
    out, bg = [], 0
    for i in range(len(steps) - 1):
        time = steps[i]
        if steps[i + 1] - bg < max_length:
            continue
        if time - bg > min_length:
            out.append([bg, time])
        bg = time
    out.append([bg, steps[-1]])
    return out
