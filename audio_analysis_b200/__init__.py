"""audio_analysis_b200 - B200-native IR decay / spectral-decay analysis.

Drop-in for the `decay`, `rt60bands`, `waterfall` and `bundle` path of
kianmcevoy/audio_analysis: the `analyse` package at the repository root re-exports
these modules under the reference's module names, so `python -m analyse.cli decay ...`
and `from analyse.decay import analyse_decay_for_channel` keep working.

Importing this package does not touch CUDA; the first analysis call loads
`csrc/libaudio_analysis_b200.so` and needs a B200.  There is no CPU fallback.
"""

from .io import (  # noqa: F401
    DEFAULT_EXPECTED_SAMPLE_RATE_HZ,
    LoadedAudio,
    convert_wav_samples_to_float32,
    downmix_to_mono,
    duplicate_mono_to_stereo,
    get_channel,
    get_left_right,
    load_wav_file,
    validate_audio_format,
)

__version__ = "0.1.0"

__all__ = [
    "LoadedAudio",
    "DEFAULT_EXPECTED_SAMPLE_RATE_HZ",
    "convert_wav_samples_to_float32",
    "downmix_to_mono",
    "duplicate_mono_to_stereo",
    "get_channel",
    "get_left_right",
    "load_wav_file",
    "validate_audio_format",
]
