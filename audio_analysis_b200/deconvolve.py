"""Deconvolution of a recorded sweep into an impulse response - mirror of the reference's `analyse/deconvolve.py`
(settings :60-75, `deconvolve_impulse_response` :124-193, `deconvolve_from_wav_files` :201-259,
`default_output_ir_path` :262-268), computed by `aa_deconvolve_batch` (csrc/deconv.cu):

    H = Y conj(X) / (|X|^2 + eps),  eps = regularization_relative * max(1e-30, max |X|^2),  n_fft = next_pow2(max(len))

fused into the first pass of the inverse FFT, then DC removal and per-file peak normalisation on the device.
`deconvolve_batch` is the batched form the reference lacks: many captures of one sweep in one call (BASELINE.json
config 4: 64 captures -> 64 impulse responses that stay on the device for rt60bands + waterfall).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import engine
from .io import convert_wav_samples_to_float32, ensure_2d_channel_array, load_wav_file


@dataclass(frozen=True)
class DeconvolveSettings:
    regularization_relative: float = 1e-10
    normalise_peak: bool = True
    target_peak: float = 0.95
    remove_dc: bool = True
    output_length_mode: str = "recorded"  # "recorded" | "full_fft"


@dataclass(frozen=True)
class DeconvolvedImpulseResponse:
    samples: np.ndarray        # shape (N, C), float32
    sample_rate_hz: int
    recorded_file_path: Path
    sweep_file_path: Path


def _next_power_of_two(n: int) -> int:
    return 1 if n <= 1 else 1 << (int(n - 1).bit_length())


def _downmix_to_mono_1d(samples_2d: np.ndarray) -> np.ndarray:
    if samples_2d.ndim != 2:
        raise ValueError("Expected a 2D array (N,C).")
    return np.mean(samples_2d.astype(np.float64, copy=False), axis=1).astype(np.float32)


def _output_length(settings: DeconvolveSettings, n_recorded: int, n_fft: int) -> int:
    if settings.output_length_mode == "recorded":
        return n_recorded
    if settings.output_length_mode == "full_fft":
        return n_fft
    raise ValueError(f"Unknown output_length_mode: {settings.output_length_mode}")


def deconvolve_batch(recorded_rows, sweep_samples_1d, settings: Optional[DeconvolveSettings] = None,
                     channels_per_file: int = 1, device=None):
    """Impulse responses of many recorded channels of equal length, all excited by the same mono sweep.

    `recorded_rows`: [rows, n_recorded] float32 array / tensor / engine.Rows, one channel per row; the channels of one
    file are `channels_per_file` consecutive rows (peak normalisation is per file, deconvolve.py:190).  Returns
    (ir rows as engine.Rows on the device, n_out)."""
    import torch

    settings = settings or DeconvolveSettings()
    rows = recorded_rows if isinstance(recorded_rows, engine.Rows) else engine.make_rows(recorded_rows, device)
    sweep = torch.as_tensor(np.asarray(sweep_samples_1d, dtype=np.float32)) if not isinstance(sweep_samples_1d, torch.Tensor) \
        else sweep_samples_1d.to(torch.float32)
    n_rec, n_sweep = int(rows.n), int(sweep.numel())
    if n_rec < 8 or n_sweep < 8:
        raise ValueError("Recorded and sweep must both contain at least a few samples.")
    n_fft = _next_power_of_two(max(n_rec, n_sweep))
    n_out = _output_length(settings, n_rec, n_fft)
    ir = engine.deconvolve(rows, sweep, n_fft, float(settings.regularization_relative), bool(settings.remove_dc),
                           float(settings.target_peak) if settings.normalise_peak else 0.0, int(channels_per_file), n_out)
    hl = np.full(rows.n_rows, n_out, dtype=np.int64)
    return engine.Rows(x=ir, lengths=None, n=n_out, host_lengths=hl), n_out


def deconvolve_impulse_response(
    recorded_samples_2d: np.ndarray,
    sweep_samples_1d: np.ndarray,
    sample_rate_hz: int,
    settings: DeconvolveSettings,
) -> np.ndarray:
    """IR for each channel of `recorded_samples_2d` (N, C) against the mono sweep; returns (N_out, C) float32
    (reference deconvolve.py:124-193)."""
    rec = ensure_2d_channel_array(convert_wav_samples_to_float32(np.asarray(recorded_samples_2d)))
    sweep = np.asarray(sweep_samples_1d, dtype=np.float32)
    if rec.shape[0] < 8 or sweep.size < 8:
        raise ValueError("Recorded and sweep must both contain at least a few samples.")
    channels = int(rec.shape[1])
    ir_rows, n_out = deconvolve_batch(np.ascontiguousarray(rec.T), sweep, settings, channels_per_file=channels)
    return np.ascontiguousarray(ir_rows.x[:, :n_out].cpu().numpy().T.astype(np.float32))


def _write_wav_float32(path: Path, sample_rate_hz: int, samples_2d: np.ndarray) -> None:
    from scipy.io import wavfile  # lazy, like io.load_wav_file

    path.parent.mkdir(parents=True, exist_ok=True)
    wavfile.write(str(path), int(sample_rate_hz), samples_2d.astype(np.float32, copy=False))


def deconvolve_from_wav_files(
    recorded_wav_file_path: str | Path,
    sweep_wav_file_path: str | Path,
    settings: Optional[DeconvolveSettings] = None,
    output_ir_wav_file_path: Optional[str | Path] = None,
) -> DeconvolvedImpulseResponse:
    """Load recorded + sweep WAVs (mono or stereo, 48 kHz; the sweep is downmixed to mono), deconvolve, optionally
    write the IR as a float32 WAV (reference deconvolve.py:201-259)."""
    settings = settings or DeconvolveSettings()
    recorded = load_wav_file(recorded_wav_file_path, expected_channel_mode="mono_or_stereo",
                             allow_mono_and_upmix_to_stereo=False)
    sweep = load_wav_file(sweep_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    if recorded.sample_rate_hz != sweep.sample_rate_hz:
        raise ValueError(f"Sample rate mismatch: recorded={recorded.sample_rate_hz} Hz, sweep={sweep.sample_rate_hz} Hz")
    samples = deconvolve_impulse_response(recorded.samples, _downmix_to_mono_1d(sweep.samples),
                                          recorded.sample_rate_hz, settings)
    ir = DeconvolvedImpulseResponse(samples=samples, sample_rate_hz=int(recorded.sample_rate_hz),
                                    recorded_file_path=Path(recorded.file_path), sweep_file_path=Path(sweep.file_path))
    if output_ir_wav_file_path is not None:
        _write_wav_float32(Path(output_ir_wav_file_path), ir.sample_rate_hz, ir.samples)
    return ir


def default_output_ir_path(recorded_wav_file_path: str | Path) -> Path:
    """<recorded_stem>_ir.wav next to the recording (reference deconvolve.py:262-268)."""
    p = Path(recorded_wav_file_path)
    return p.with_name(f"{p.stem}_ir.wav")
