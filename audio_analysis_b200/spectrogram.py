"""Spectrogram (dense STFT magnitude in dB) on the B200 path - SURVEY.md 8f N2.

Drop-in mirror of the reference's `analyse/spectrogram.py` API.  It is the waterfall's STFT
(`analyse/spectrogram.py:107-160` is the same loop as `analyse/waterfall.py:188-230`) with every
frame kept: time selection on the host/device bookkeeping level (trim to peak, ignore leading,
duration; reference :176-194), then `aa_stft_db_batch` computes all frames of all channels - a run of
overlapping frames is staged by one TMA bulk copy, so each sample is read from HBM once.
The reference stores `magnitude_db` as (F, T); the device writes frame-major and the host hands out
the transposed view.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from . import engine
from .io import get_analysis_channels, load_wav_file

_TOO_SHORT = "Not enough samples after trimming/selection for spectrogram (need at least n_fft)."
_DENSE_NFFT = (1024, 2048, 4096, 8192)


@dataclass(frozen=True)
class SpectrogramAnalysisSettings:
    use_mono_downmix_for_stereo: bool = False
    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    analysis_duration_seconds: Optional[float] = None
    n_fft: int = 4096
    hop_length: int = 512
    use_hann_window: bool = True
    floor_db: float = -120.0
    f_min_hz: float = 20.0
    f_max_hz: float = 20000.0
    dynamic_range_db: Optional[float] = 90.0  # None: percentile colour scale


@dataclass(frozen=True)
class SpectrogramPlotSettings:
    vmin_db: Optional[float] = None
    vmax_db: Optional[float] = None


@dataclass(frozen=True)
class ChannelSpectrogramResult:
    channel_name: str
    sample_rate_hz: int
    analysis_start_sample_index: int
    analysis_length_samples: int
    time_seconds: np.ndarray      # (T,) float32, frame start times
    frequency_hz: np.ndarray      # (F,) float32
    magnitude_db: np.ndarray      # (F, T) float32


# --------------------------------------------------------------------------------------
# batched analysis
# --------------------------------------------------------------------------------------


def run_spectrogram(rows: engine.Rows, sample_rate_hz: int, settings: SpectrogramAnalysisSettings, peaks=None):
    """Device-level call: (mag_db [B, T_max, F] float32 frame-major, start [B] int32, span [B] int32, T_max).

    `peaks` (device int32 [B]) may come from the decay kernel.  Rows shorter than n_fft after the time selection
    get zero frames; the per-channel wrappers turn that into the reference's ValueError.
    """
    import torch

    n_fft, hop = int(settings.n_fft), int(settings.hop_length)
    if n_fft <= 0 or hop <= 0:
        raise ValueError("n_fft and hop_length must be positive.")
    if n_fft & (n_fft - 1) or not 64 <= n_fft <= 8192:
        raise ValueError("this build computes STFTs for power-of-two n_fft in [64, 8192]")
    dev = rows.x.device
    b = rows.n_rows
    lengths = rows.lengths if rows.lengths is not None else torch.full((b,), rows.n, dtype=torch.int32, device=dev)
    if settings.trim_to_peak:   # reference spectrogram.py:176-179
        start = (peaks if peaks is not None else engine.peak_index(rows)).to(torch.int32)
    else:
        start = torch.zeros(b, dtype=torch.int32, device=dev)
    span = lengths - start
    if settings.ignore_leading_seconds > 0.0:   # :181-185
        skip = int(round(float(settings.ignore_leading_seconds) * float(sample_rate_hz)))
        skip_t = torch.clamp(torch.full_like(span, max(0, skip)), max=span)
        start = start + skip_t
        span = span - skip_t
    if settings.analysis_duration_seconds is not None:   # :187-190
        keep = int(round(float(settings.analysis_duration_seconds) * float(sample_rate_hz)))
        span = torch.clamp(torch.full_like(span, max(0, keep)), max=span)
    start = start.contiguous()
    span = span.to(torch.int32).contiguous()
    t_max = max(0, 1 + (rows.n - n_fft) // hop) if rows.n >= n_fft else 0
    if n_fft in _DENSE_NFFT:
        mag = engine.stft_db(rows, start, span, n_fft, hop, settings.use_hann_window, settings.floor_db, t_max)
    else:
        mag = _stft_via_slices(rows, start, span, settings, t_max)
    return mag, start, span, t_max


def _stft_via_slices(rows: engine.Rows, start, span, settings: SpectrogramAnalysisSettings, t_max: int):
    """Small n_fft (64 ... 512): the slice kernel with no reference level, 64 frames per call."""
    import torch

    from . import _native as nat

    n_fft, hop = int(settings.n_fft), int(settings.hop_length)
    dev = rows.x.device
    b = rows.n_rows
    f = n_fft // 2 + 1
    out = torch.zeros((b, t_max, f), dtype=torch.float32, device=dev)
    n_frames = torch.where(span >= n_fft, 1 + torch.div(span - n_fft, hop, rounding_mode="floor"),
                           torch.zeros_like(span))
    params = engine.csd_params(n_fft, hop, settings.use_hann_window, 0, f, "global_max", settings.floor_db, 1.0,
                               raw_db=True)
    for j0 in range(0, t_max, nat.AA_MAX_SLICES):
        cnt = min(nat.AA_MAX_SLICES, t_max - j0)
        frames = (torch.arange(cnt, dtype=torch.int32, device=dev) + j0).unsqueeze(0).expand(b, cnt).contiguous()
        have = torch.clamp(n_frames - j0, min=0, max=cnt).to(torch.int32).contiguous()
        part = engine.csd(rows, start, frames, have, params)
        valid = torch.arange(cnt, device=dev).unsqueeze(0) < have.unsqueeze(1)
        out[:, j0:j0 + cnt] = torch.where(valid.unsqueeze(2), part, torch.zeros_like(part))
    return out


def analyse_spectrogram_batch(channels, sample_rate_hz: int, channel_names: Optional[Sequence[str]] = None,
                              settings: Optional[SpectrogramAnalysisSettings] = None, device=None,
                              peaks=None) -> List[ChannelSpectrogramResult]:
    settings = settings or SpectrogramAnalysisSettings()
    rows = channels if isinstance(channels, engine.Rows) else engine.make_rows(channels, device)
    names = list(channel_names) if channel_names is not None else [f"ch{i}" for i in range(rows.n_rows)]
    mag, start, span, _ = run_spectrogram(rows, sample_rate_hz, settings, peaks)
    n_fft, hop = int(settings.n_fft), int(settings.hop_length)
    start_h = start.cpu().numpy()
    span_h = span.cpu().numpy()
    if np.any(span_h < n_fft):
        raise ValueError(_TOO_SHORT)
    table = mag.cpu().numpy()
    freq = np.fft.rfftfreq(n_fft, d=1.0 / float(sample_rate_hz)).astype(np.float32)
    results = []
    for b in range(rows.n_rows):
        t = 1 + (int(span_h[b]) - n_fft) // hop
        times = (np.arange(t, dtype=np.float32) * float(hop) / float(sample_rate_hz)).astype(np.float32)
        results.append(ChannelSpectrogramResult(
            channel_name=str(names[b]),
            sample_rate_hz=int(sample_rate_hz),
            analysis_start_sample_index=int(start_h[b]),
            analysis_length_samples=int(span_h[b]),
            time_seconds=times,
            frequency_hz=freq,
            magnitude_db=table[b, :t, :].T,    # (F, T) like the reference (a view of the frame-major table)
        ))
    return results


def analyse_spectrogram_for_channel(
    samples: np.ndarray,
    sample_rate_hz: int,
    channel_name: str,
    settings: SpectrogramAnalysisSettings,
) -> ChannelSpectrogramResult:
    """One channel (reference spectrogram.py:163-216)."""
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("analyse_spectrogram_for_channel expects a 1D mono array.")
    if samples.size < int(settings.n_fft):
        raise ValueError(_TOO_SHORT)
    return analyse_spectrogram_batch(samples[None, :], sample_rate_hz, [channel_name], settings)[0]


def analyse_spectrogram_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[SpectrogramAnalysisSettings] = None,
) -> List[ChannelSpectrogramResult]:
    """Reference spectrogram.py:219-247; the channels of one file share a batch."""
    settings = settings or SpectrogramAnalysisSettings()
    audio = load_wav_file(input_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    chans = get_analysis_channels(audio, use_mono_downmix_for_stereo=settings.use_mono_downmix_for_stereo)
    data = np.stack([c for _, c in chans], axis=0)
    return analyse_spectrogram_batch(data, audio.sample_rate_hz, [n for n, _ in chans], settings)


# --------------------------------------------------------------------------------------
# figures (optional) and text
# --------------------------------------------------------------------------------------


def _cell_edges(centres: np.ndarray, lone_width: float) -> np.ndarray:
    """Cell boundaries half-way between neighbouring centres (what pcolormesh wants)."""
    c = centres.astype(np.float64)
    if c.size == 1:
        return np.array([c[0], c[0] + lone_width])
    mid = 0.5 * (c[:-1] + c[1:])
    return np.concatenate(([c[0] - 0.5 * (c[1] - c[0])], mid, [c[-1] + 0.5 * (c[-1] - c[-2])]))


def plot_spectrogram_figure(
    result: ChannelSpectrogramResult,
    analysis_settings: SpectrogramAnalysisSettings,
    plot_settings: SpectrogramPlotSettings,
    title: Optional[str] = None,
):
    """Log-frequency pcolormesh with a colour bar (same content as reference spectrogram.py:255-338)."""
    from .plotting import create_figure_and_axis, pyplot

    pyplot(optional=False)
    import matplotlib.ticker as mticker

    figure, axis = create_figure_and_axis(title=title)
    nyquist = 0.5 * float(result.sample_rate_hz)
    f_lo = float(np.clip(analysis_settings.f_min_hz, 1.0, nyquist))
    f_hi = float(np.clip(analysis_settings.f_max_hz, f_lo, nyquist))
    keep = (result.frequency_hz >= f_lo) & (result.frequency_hz <= f_hi)
    mag = result.magnitude_db[keep, :]
    if mag.size == 0:
        raise ValueError("Spectrogram frequency selection is empty (check f_min_hz/f_max_hz).")
    vmax = float(plot_settings.vmax_db) if plot_settings.vmax_db is not None else float(np.percentile(mag, 99.5))
    if plot_settings.vmin_db is not None:
        vmin = float(plot_settings.vmin_db)
    elif analysis_settings.dynamic_range_db is not None:
        vmin = vmax - float(analysis_settings.dynamic_range_db)
    else:
        vmin = float(np.percentile(mag, 5.0))
    t_edges = _cell_edges(result.time_seconds, 1e-3)
    f_edges = np.maximum(_cell_edges(result.frequency_hz[keep], 1.0), 1e-6)
    mesh = axis.pcolormesh(t_edges, f_edges, mag, shading="auto", vmin=vmin, vmax=vmax)
    axis.set_xlabel("Time (s)")
    axis.set_ylabel("Frequency (Hz)")
    axis.set_yscale("log")
    axis.set_ylim(f_lo, f_hi)
    axis.set_yticks([20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000])
    axis.yaxis.set_major_formatter(mticker.FuncFormatter(
        lambda x, pos: f"{int(x / 1000)}k" if x >= 1000.0 else f"{int(x)}"))
    axis.yaxis.set_minor_formatter(mticker.NullFormatter())
    axis.grid(True, which="both", linestyle=":", linewidth=0.5)
    figure.colorbar(mesh, ax=axis, label="Magnitude (dB)")
    return figure


def plot_spectrogram_from_wav_file(
    input_wav_file_path: str | Path,
    analysis_settings: Optional[SpectrogramAnalysisSettings] = None,
    plot_settings: Optional[SpectrogramPlotSettings] = None,
    output_basename: Optional[str | Path] = None,
    show_interactive: bool = True,
) -> List[ChannelSpectrogramResult]:
    """Analyse, then one `<basename>_spectrogram_<CH>.png` per channel when matplotlib is available (:341-389)."""
    from .plotting import finalize_and_show_or_save, pyplot

    analysis_settings = analysis_settings or SpectrogramAnalysisSettings()
    plot_settings = plot_settings or SpectrogramPlotSettings()
    results = analyse_spectrogram_from_wav_file(input_wav_file_path, analysis_settings)
    if (output_basename is not None or show_interactive) and pyplot() is not None:
        for r in results:
            figure = plot_spectrogram_figure(r, analysis_settings, plot_settings,
                                             title=f"Spectrogram — {input_wav_file_path} — {r.channel_name}")
            target = None
            if output_basename is not None:
                base = Path(output_basename)
                target = base.with_name(f"{base.stem}_spectrogram_{r.channel_name}.png").with_suffix(".png")
            finalize_and_show_or_save(figure, output_path=target, show_interactive=show_interactive)
    return results


def summarise_spectrogram_results_text(results: List[ChannelSpectrogramResult]) -> str:
    """`[ch] start_sample= len_samples= dur= stft(n_fft=, frames=)` per channel (reference :397-406)."""
    return "\n".join(
        f"[{r.channel_name}] start_sample={r.analysis_start_sample_index}  "
        f"len_samples={r.analysis_length_samples}  "
        f"dur={float(r.analysis_length_samples) / float(r.sample_rate_hz):.3f}s  "
        f"stft(n_fft={r.magnitude_db.shape[0] * 2 - 2}, frames={r.magnitude_db.shape[1]})"
        for r in results
    )
