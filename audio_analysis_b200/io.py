"""WAV input and channel policy (host side).

Mirror of the reference's `analyse/io.py` public surface: same names, argument
meaning and error behaviour.  This stays on the CPU: it is file decoding, the
analysed channels are handed to the device engine afterwards.  The batched
bundle path decodes PCM16 taps on the GPU instead (`engine.pcm16_to_rows`).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List, Literal, Tuple

import numpy as np

ChannelMode = Literal["mono", "stereo", "mono_or_stereo"]
DEFAULT_EXPECTED_SAMPLE_RATE_HZ = 48_000


@dataclass(frozen=True)
class LoadedAudio:
    """Audio in the internal representation: float32 in [-1, 1], shape (num_samples, num_channels)."""

    samples: np.ndarray
    sample_rate_hz: int
    file_path: Path


_PCM_SCALE = {np.dtype(np.int16): 32768.0, np.dtype(np.int32): 2147483648.0}


def convert_wav_samples_to_float32(samples_from_wav: np.ndarray) -> np.ndarray:
    """int16 / int32 PCM or float WAV payload -> float32 clipped to [-1, 1] (reference io.py:46-64, :98-113)."""
    dt = samples_from_wav.dtype
    if np.issubdtype(dt, np.floating):
        as_float = samples_from_wav.astype(np.float32, copy=False)
    elif np.issubdtype(dt, np.integer):
        scale = _PCM_SCALE.get(np.dtype(dt))
        if scale is None:
            raise ValueError(f"Unsupported integer PCM dtype: {dt}")
        as_float = samples_from_wav.astype(np.float32) / scale
    else:
        raise ValueError(f"Unsupported WAV dtype: {dt}")
    return np.clip(as_float, -1.0, 1.0).astype(np.float32)


def ensure_2d_channel_array(float_samples: np.ndarray) -> np.ndarray:
    """(N,) -> (N, 1); (N, C) unchanged (reference io.py:116-126)."""
    if float_samples.ndim == 2:
        return float_samples
    if float_samples.ndim == 1:
        return float_samples.reshape((-1, 1))
    raise ValueError(f"Expected 1D or 2D audio array, got shape {float_samples.shape}")


def duplicate_mono_to_stereo(float_samples: np.ndarray) -> np.ndarray:
    """Mono -> two identical channels; stereo passes through (reference io.py:129-143)."""
    arr = ensure_2d_channel_array(float_samples)
    count = arr.shape[1]
    if count == 2:
        return arr.astype(np.float32)
    if count == 1:
        return np.repeat(arr, 2, axis=1).astype(np.float32)
    raise ValueError(f"Expected mono or stereo for upmix, got {count} channels")


def downmix_to_mono(float_samples: np.ndarray) -> np.ndarray:
    """Channel average, shape (N, 1) (reference io.py:146-154)."""
    arr = ensure_2d_channel_array(float_samples)
    return np.mean(arr, axis=1, dtype=np.float32).reshape((-1, 1)).astype(np.float32)


def validate_audio_format(
    loaded_audio: LoadedAudio,
    expected_sample_rate_hz: int = DEFAULT_EXPECTED_SAMPLE_RATE_HZ,
    expected_channel_mode: ChannelMode = "stereo",
) -> None:
    """Sample-rate and channel-count checks with the reference's messages (io.py:157-178)."""
    if loaded_audio.sample_rate_hz != expected_sample_rate_hz:
        raise ValueError(
            f"Expected sample rate {expected_sample_rate_hz} Hz, "
            f"but got {loaded_audio.sample_rate_hz} Hz for file {loaded_audio.file_path}"
        )
    count = loaded_audio.samples.shape[1]
    where = f"for file {loaded_audio.file_path}"
    if expected_channel_mode == "mono" and count != 1:
        raise ValueError(f"Expected mono (1 channel) but got {count} channels {where}")
    if expected_channel_mode == "stereo" and count != 2:
        raise ValueError(f"Expected stereo (2 channels) but got {count} channels {where}")
    if expected_channel_mode == "mono_or_stereo" and count not in (1, 2):
        raise ValueError(f"Expected mono or stereo (1 or 2 channels) but got {count} channels {where}")


def load_wav_file(
    wav_file_path: str | Path,
    expected_sample_rate_hz: int = DEFAULT_EXPECTED_SAMPLE_RATE_HZ,
    expected_channel_mode: ChannelMode = "stereo",
    allow_mono_and_upmix_to_stereo: bool = True,
) -> LoadedAudio:
    """Read a WAV, convert to float32 (N, C), optionally upmix mono, validate (reference io.py:181-221)."""
    from scipy.io import wavfile  # lazy: only file-based entry points need scipy

    path = Path(wav_file_path)
    rate, payload = wavfile.read(str(path))
    samples = ensure_2d_channel_array(convert_wav_samples_to_float32(payload))
    if expected_channel_mode == "stereo" and allow_mono_and_upmix_to_stereo and samples.shape[1] == 1:
        samples = duplicate_mono_to_stereo(samples)
    loaded = LoadedAudio(samples=samples.astype(np.float32, copy=False), sample_rate_hz=int(rate), file_path=path)
    validate_audio_format(loaded, expected_sample_rate_hz, expected_channel_mode)
    return loaded


def get_analysis_channels(
    loaded_audio: LoadedAudio,
    use_mono_downmix_for_stereo: bool = False,
) -> List[Tuple[str, np.ndarray]]:
    """[(name, 1-D samples)]: mono | left,right | mono = 0.5*(L+R) (reference io.py:66-95)."""
    count = loaded_audio.samples.shape[1]
    if count == 1:
        return [("mono", loaded_audio.samples[:, 0].astype(np.float32, copy=False))]
    if count == 2:
        left = loaded_audio.samples[:, 0].astype(np.float32, copy=False)
        right = loaded_audio.samples[:, 1].astype(np.float32, copy=False)
        if use_mono_downmix_for_stereo:
            return [("mono", 0.5 * (left + right))]
        return [("left", left), ("right", right)]
    raise ValueError(f"Unsupported channel count: {count}")


def get_channel(loaded_audio: LoadedAudio, channel_index: int) -> np.ndarray:
    """One channel as a 1-D float32 array (reference io.py:224-235)."""
    count = loaded_audio.samples.shape[1]
    if not 0 <= channel_index < count:
        raise ValueError(f"channel_index out of range: {channel_index} for {count} channels")
    return loaded_audio.samples[:, channel_index].astype(np.float32, copy=False)


def get_left_right(loaded_audio: LoadedAudio) -> Tuple[np.ndarray, np.ndarray]:
    """(left, right) of a stereo file (reference io.py:238-247)."""
    validate_audio_format(loaded_audio, expected_channel_mode="stereo")
    return get_channel(loaded_audio, 0), get_channel(loaded_audio, 1)
