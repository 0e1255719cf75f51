"""The bundle wire format: 16-bit PCM WAV taps as the reference's C++ AnalysisRecorder writes them
(include/analysis/recorder.hpp:55-90: a 44-byte RIFF header - "fmt " chunk of 16 bytes, audio format 1, 16 bits,
interleaved channels - followed by the "data" chunk).

The reference loads every tap with scipy, converts it to float32 on the CPU and does so again for every report block
(analyse/report.py:193-199, :237-386).  Here the payload bytes of all taps are read ONCE, straight into one pinned
int16 staging buffer; the device decodes them (`aa_pcm16_to_float_batch`: sample / 32768, channel split or 0.5 (L + R),
io.py:46-95), so half the bytes cross PCIe and the host never touches a sample.  Files in any other WAV flavour
(int32, float, extra chunks before "fmt ") take the scipy loader of io.py.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class Pcm16Info:
    sample_rate_hz: int
    channels: int
    frames: int
    data_offset: int


def probe_pcm16(path: str | Path) -> Optional[Pcm16Info]:
    """Header of a PCM16 WAV, or None when the file is anything else (then use io.load_wav_file)."""
    try:
        with open(path, "rb") as f:
            head = f.read(12)
            if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
                return None
            fmt = None
            while True:
                ck = f.read(8)
                if len(ck) < 8:
                    return None
                tag, size = ck[:4], struct.unpack("<I", ck[4:])[0]
                if tag == b"fmt ":
                    body = f.read(size + (size & 1))
                    if size < 16:
                        return None
                    audio_fmt, chans, rate, _, block_align, bits = struct.unpack("<HHIIHH", body[:16])
                    if audio_fmt != 1 or bits != 16 or chans < 1 or block_align != 2 * chans:
                        return None
                    fmt = (int(rate), int(chans))
                elif tag == b"data":
                    if fmt is None:
                        return None
                    return Pcm16Info(fmt[0], fmt[1], int(size) // (2 * fmt[1]), f.tell())
                else:
                    f.seek(size + (size & 1), 1)
    except OSError:
        return None


@dataclass
class StagedTaps:
    """All taps of a batch in one pinned int16 buffer [files, frames * channels] (interleaved as on disk)."""

    pcm: "object"            # torch.Tensor, int16, pinned when a CUDA device is present
    sample_rate_hz: int
    channels: int
    frames: int


_staging_cache = {}   # (files, width, pinned) -> tensor: pinning hundreds of MB costs more than reading them


def _staging_buffer(files: int, width: int, pin: bool):
    import torch

    key = (int(files), int(width), bool(pin))
    buf = _staging_cache.get(key)
    if buf is None:
        _staging_cache.clear()          # one live buffer: bundles are processed one after the other
        try:
            buf = torch.empty((files, width), dtype=torch.int16, pin_memory=bool(pin))
        except RuntimeError:
            buf = torch.empty((files, width), dtype=torch.int16)
        _staging_cache[key] = buf
    return buf


def stage_pcm16(paths: Sequence[str | Path], pin: bool = True) -> Optional[StagedTaps]:
    """Read the payloads of PCM16 WAV files of one common shape into a pinned staging buffer (reused between calls: the
    returned tensor is valid until the next call).  None when any file is not PCM16 or the files differ in sample rate,
    channel count or length (the caller then loads them one by one).  Files are read by a small thread pool."""
    from concurrent.futures import ThreadPoolExecutor

    import torch

    infos: List[Pcm16Info] = []
    for p in paths:
        info = probe_pcm16(p)
        if info is None or info.frames < 1:
            return None
        infos.append(info)
    if not infos:
        return None
    first = infos[0]
    if any((i.sample_rate_hz, i.channels, i.frames) != (first.sample_rate_hz, first.channels, first.frames) for i in infos):
        return None
    width = first.frames * first.channels
    buf = _staging_buffer(len(infos), width, bool(pin and torch.cuda.is_available()))
    view = buf.numpy()

    def read_one(row: int) -> None:
        with open(paths[row], "rb") as f:
            f.seek(infos[row].data_offset)
            got = f.readinto(memoryview(view[row]).cast("B"))
        if got != 2 * width:
            raise ValueError(f"{paths[row]}: data chunk shorter than its header says")

    if len(infos) >= 8:
        with ThreadPoolExecutor(max_workers=8) as pool:
            list(pool.map(read_one, range(len(infos))))
    else:
        for row in range(len(infos)):
            read_one(row)
    return StagedTaps(buf, first.sample_rate_hz, first.channels, first.frames)


def write_pcm16_wav(path: str | Path, pcm: np.ndarray, sample_rate_hz: int) -> None:
    """[frames, channels] (or [frames]) int16 -> the recorder's file layout (recorder.hpp:55-90).  Test / bench helper."""
    arr = np.ascontiguousarray(pcm, dtype="<i2")
    chans = 1 if arr.ndim == 1 else int(arr.shape[1])
    data = arr.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt ")
        f.write(struct.pack("<IHHIIHH", 16, 1, chans, int(sample_rate_hz), int(sample_rate_hz) * 2 * chans, 2 * chans, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)
