"""Audio decode for the step before the frontend: WAV / raw PCM bytes -> mono fp32 waveform at the model's rate,
on the GPU.  The reference's loader (funasr/utils/load_utils.py:48-179: torchaudio.load / ffmpeg / librosa, channel average,
torchaudio Resample when the rates differ) does this on the CPU; here only the RIFF header is parsed on the host — the sample bytes
go to the device as they are and are converted (fa_pcm_decode), mixed down and resampled (fa_resample) there."""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np
import torch

from . import _abi
from .resample import resample

_FMT = {"f32": 0, "s16": 1, "s24": 2, "s32": 3, "u8": 4}
_BYTES = {0: 4, 1: 2, 2: 3, 3: 4, 4: 1}


def parse_wav_header(data: bytes) -> Tuple[int, int, int, int, int]:
    """RIFF/WAVE -> (sample_format code, channels, sample_rate, data_offset, data_bytes).  PCM (tag 1: 8/16/24/32 bit), IEEE float
    (tag 3: 32 bit) and WAVE_FORMAT_EXTENSIBLE with those sub-formats."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise _abi.FunasrB200Error("not a RIFF/WAVE file")
    pos, fmt = 12, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = pos + 8
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", data[body:body + 16])
            if tag == 0xFFFE and size >= 40:                                   # extensible: the sub-format GUID starts with the real tag
                tag = struct.unpack("<H", data[body + 24:body + 26])[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            if fmt is None:
                raise _abi.FunasrB200Error("WAV: data chunk before fmt chunk")
            tag, ch, rate, bits = fmt
            code = {(1, 8): 4, (1, 16): 1, (1, 24): 2, (1, 32): 3, (3, 32): 0}.get((tag, bits))
            if code is None:
                raise _abi.FunasrB200Error("WAV: unsupported sample format (tag %d, %d bits)" % (tag, bits))
            return code, ch, rate, body, min(size, len(data) - body)
        pos = body + size + (size & 1)
    raise _abi.FunasrB200Error("WAV: no data chunk")


def decode_pcm(pcm: bytes, sample_format: str, channels: int, device) -> torch.Tensor:
    """Raw interleaved PCM bytes -> mono fp32 [frames] on the device."""
    code = _FMT[sample_format] if isinstance(sample_format, str) else int(sample_format)
    frames = len(pcm) // (_BYTES[code] * channels)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _abi.FunasrB200Error("funasr_b200.audio needs a CUDA device (no CPU path)")
    raw = torch.frombuffer(bytearray(pcm[: frames * _BYTES[code] * channels]), dtype=torch.uint8).to(dev, non_blocking=True) if frames else \
        torch.zeros(0, dtype=torch.uint8, device=dev)
    out = torch.empty(frames, dtype=torch.float32, device=dev)
    _abi.check(_abi.load().fa_pcm_decode(raw.data_ptr(), code, channels, frames, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "fa_pcm_decode")
    return out


def load_audio(data, fs: int = 16000, device="cuda", audio_fs: int = None, sample_format: str = "s16", channels: int = 1) -> torch.Tensor:
    """WAV bytes / a .wav or .pcm path / raw PCM bytes -> mono fp32 waveform at `fs` on the device.  Raw PCM uses `audio_fs`
    (default 16000), `sample_format` and `channels`; WAV carries its own."""
    if isinstance(data, str):
        with open(data, "rb") as f:
            data = f.read()
    if data[:4] == b"RIFF":
        code, ch, rate, off, nbytes = parse_wav_header(data)
        wav = decode_pcm(data[off:off + nbytes], code, ch, device)
    else:
        rate = int(audio_fs or 16000)
        wav = decode_pcm(data, sample_format, channels, device)
    if rate != fs and wav.numel():
        out, lens = resample(wav[None], torch.tensor([wav.numel()], dtype=torch.int32), rate, fs)
        wav = out[0, : int(lens[0])].contiguous()
    return wav
