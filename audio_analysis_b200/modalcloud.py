"""Modal cloud: per log-frequency bin RT60 estimates from the STFT - SURVEY.md 8f N3.

Drop-in mirror of the reference's `analyse/modalcloud.py` API.  Device side: the dense STFT of every
channel (`aa_stft_db_batch`, n_fft 8192 by default), then `aa_modal_cloud_batch` averages the linear
magnitudes of the rfft lines inside every log bin (reference :176-207), normalises each bin's curve to
its own peak, finds the two range crossings and fits the line (reference :215-281, :339-371).  Host
side: the log-bin edges and their rfft-line ranges (a function of the settings only, :168-173, :324-333)
and the conversion of the [bins, 2] table into `ModalPoint`s.
"""

from __future__ import annotations

from dataclasses import dataclass
from math import ceil, log2
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from .io import get_analysis_channels, load_wav_file
from .spectrogram import SpectrogramAnalysisSettings, run_spectrogram

_TOO_SHORT = "Not enough samples after trimming/selection for modal cloud (need at least n_fft)."


@dataclass(frozen=True)
class ModalCloudAnalysisSettings:
    use_mono_downmix_for_stereo: bool = False
    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    analysis_duration_seconds: Optional[float] = None
    n_fft: int = 8192
    hop_length: int = 512
    use_hann_window: bool = True
    f_min_hz: float = 20.0
    f_max_hz: float = 20000.0
    log_bins_per_octave: int = 24
    min_bins: int = 24
    floor_db: float = -120.0
    fit_lower_limit_db: float = -80.0
    t30_range_db: Tuple[float, float] = (-5.0, -35.0)
    t20_range_db: Tuple[float, float] = (-5.0, -25.0)
    edt_range_db: Tuple[float, float] = (0.0, -10.0)
    metric: str = "t30"  # "t30" | "t20" | "edt"
    min_fit_points: int = 10
    min_peak_db_above_floor: float = 20.0


@dataclass(frozen=True)
class ModalCloudPlotSettings:
    secondary_channel_alpha: float = 0.7
    show_median_curve: bool = True
    median_octave_window: float = 0.25
    ylim_seconds: Optional[Tuple[float, float]] = None


@dataclass(frozen=True)
class ModalPoint:
    centre_hz: float
    rt60_seconds: float
    r_squared: float


@dataclass(frozen=True)
class ChannelModalCloudResult:
    channel_name: str
    sample_rate_hz: int
    analysis_start_sample_index: int
    analysis_length_samples: int
    metric: str
    points: List[ModalPoint]


# --------------------------------------------------------------------------------------
# host-side tables
# --------------------------------------------------------------------------------------


def _build_log_bins(f_min_hz: float, f_max_hz: float, bins_per_octave: int, min_bins: int) -> np.ndarray:
    """Log-spaced bin edges, float32 (reference modalcloud.py:168-173)."""
    lo = float(max(1.0, f_min_hz))
    hi = float(max(lo * 1.001, f_max_hz))
    octaves = float(log2(hi / lo))
    count = int(max(min_bins, ceil(octaves * float(max(4, bins_per_octave)))))
    return (lo * (2.0 ** np.linspace(0.0, octaves, count + 1, dtype=np.float64))).astype(np.float32)


def metric_and_range(settings: ModalCloudAnalysisSettings) -> Tuple[str, Tuple[float, float]]:
    """Reference :335-342: anything that is not "t20" / "edt" means "t30"."""
    m = str(settings.metric).lower()
    if m == "t20":
        return "t20", settings.t20_range_db
    if m == "edt":
        return "edt", settings.edt_range_db
    return "t30", settings.t30_range_db


def log_bin_table(settings: ModalCloudAnalysisSettings, sample_rate_hz: int):
    """(centres f32 [B], ranges int32 [B, 2]): rfft-line range [k0, k1) of every log bin.

    The reference first keeps the lines with f_min <= f <= f_max (:324-327) and then, per log bin, those with
    lo <= f < hi (:196-198); both are contiguous in k, so each bin is one range (possibly empty).
    """
    nyquist = 0.5 * float(sample_rate_hz)
    f_lo = float(np.clip(settings.f_min_hz, 1.0, nyquist))
    f_hi = float(np.clip(settings.f_max_hz, f_lo, nyquist))
    freq = np.fft.rfftfreq(int(settings.n_fft), d=1.0 / float(sample_rate_hz)).astype(np.float32)
    keep = (freq >= f_lo) & (freq <= f_hi)
    edges = _build_log_bins(f_lo, f_hi, int(settings.log_bins_per_octave), int(settings.min_bins))
    e64 = edges.astype(np.float64)
    centres = np.sqrt(e64[:-1] * e64[1:]).astype(np.float32)
    ranges = np.zeros((centres.size, 2), dtype=np.int32)
    for b in range(centres.size):
        inside = np.nonzero(keep & (freq >= float(e64[b])) & (freq < float(e64[b + 1])))[0]
        if inside.size:
            ranges[b] = (int(inside[0]), int(inside[-1]) + 1)
    return centres, ranges


# --------------------------------------------------------------------------------------
# batched analysis
# --------------------------------------------------------------------------------------


def run_modal_cloud(rows: engine.Rows, sample_rate_hz: int, settings: ModalCloudAnalysisSettings, peaks=None):
    """Device-level call: (table [B, n_bins, 2] float64 (rt60, r^2; NaN = no point), centres, start, span)."""
    import torch

    stft = SpectrogramAnalysisSettings(
        trim_to_peak=settings.trim_to_peak, ignore_leading_seconds=settings.ignore_leading_seconds,
        analysis_duration_seconds=settings.analysis_duration_seconds, n_fft=int(settings.n_fft),
        hop_length=int(settings.hop_length), use_hann_window=settings.use_hann_window, floor_db=settings.floor_db)
    mag, start, span, t_max = run_spectrogram(rows, sample_rate_hz, stft, peaks)
    n_fft, hop = int(settings.n_fft), int(settings.hop_length)
    n_frames = torch.where(span >= n_fft, 1 + torch.div(span - n_fft, hop, rounding_mode="floor"),
                           torch.zeros_like(span)).to(torch.int32).contiguous()
    centres, ranges = log_bin_table(settings, sample_rate_hz)
    _, range_db = metric_and_range(settings)
    if float(range_db[1]) > float(range_db[0]):
        raise ValueError("range_db should be (higher_db, lower_db), e.g. (-5, -35).")
    table = engine.modal_cloud(mag, n_frames, torch.from_numpy(ranges).to(mag.device), hop, sample_rate_hz, range_db,
                               settings.fit_lower_limit_db, settings.min_fit_points, settings.floor_db,
                               settings.min_peak_db_above_floor)
    return table, centres, start, span


def analyse_modal_cloud_batch(channels, sample_rate_hz: int, channel_names: Optional[Sequence[str]] = None,
                              settings: Optional[ModalCloudAnalysisSettings] = None, device=None,
                              peaks=None) -> List[ChannelModalCloudResult]:
    settings = settings or ModalCloudAnalysisSettings()
    rows = channels if isinstance(channels, engine.Rows) else engine.make_rows(channels, device)
    names = list(channel_names) if channel_names is not None else [f"ch{i}" for i in range(rows.n_rows)]
    table, centres, start, span = run_modal_cloud(rows, sample_rate_hz, settings, peaks)
    start_h, span_h = start.cpu().numpy(), span.cpu().numpy()
    if np.any(span_h < int(settings.n_fft)):
        raise ValueError(_TOO_SHORT)
    tab = table.cpu().numpy()
    metric, _ = metric_and_range(settings)
    results = []
    for b in range(rows.n_rows):
        pts = [ModalPoint(centre_hz=float(centres[k]), rt60_seconds=float(tab[b, k, 0]), r_squared=float(tab[b, k, 1]))
               for k in range(centres.size) if not np.isnan(tab[b, k, 0])]
        pts.sort(key=lambda p: p.centre_hz)
        results.append(ChannelModalCloudResult(
            channel_name=str(names[b]), sample_rate_hz=int(sample_rate_hz),
            analysis_start_sample_index=int(start_h[b]), analysis_length_samples=int(span_h[b]),
            metric=metric, points=pts))
    return results


def analyse_modal_cloud_for_channel(
    samples: np.ndarray,
    sample_rate_hz: int,
    channel_name: str,
    settings: ModalCloudAnalysisSettings,
) -> ChannelModalCloudResult:
    """One channel (reference modalcloud.py:284-391)."""
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("_compute_stft_magnitude_db expects a 1D mono array.")
    if samples.size < int(settings.n_fft):
        raise ValueError(_TOO_SHORT)
    return analyse_modal_cloud_batch(samples[None, :], sample_rate_hz, [channel_name], settings)[0]


def analyse_modal_cloud_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[ModalCloudAnalysisSettings] = None,
) -> List[ChannelModalCloudResult]:
    """Reference modalcloud.py:394-421; the channels of one file share a batch."""
    settings = settings or ModalCloudAnalysisSettings()
    audio = load_wav_file(input_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    chans = get_analysis_channels(audio, use_mono_downmix_for_stereo=settings.use_mono_downmix_for_stereo)
    data = np.stack([c for _, c in chans], axis=0)
    return analyse_modal_cloud_batch(data, audio.sample_rate_hz, [n for n, _ in chans], settings)


# --------------------------------------------------------------------------------------
# figures (optional) and text
# --------------------------------------------------------------------------------------


def _median_curve(points: List[ModalPoint], window_octaves: float):
    """Running median of RT60 over a window in octaves (reference :442-466): None below 8 points / 4 medians."""
    if len(points) < 8:
        return None
    half = 0.5 * float(max(0.01, window_octaves))
    f = np.array([p.centre_hz for p in points], dtype=np.float64)
    y = np.array([p.rt60_seconds for p in points], dtype=np.float64)
    lf = np.log2(f)
    xs, ys = [], []
    for i in range(f.size):
        near = (lf >= lf[i] - half) & (lf <= lf[i] + half)
        if int(near.sum()) >= 3:
            xs.append(f[i])
            ys.append(float(np.median(y[near])))
    if len(xs) < 4:
        return None
    return np.array(xs, dtype=np.float32), np.array(ys, dtype=np.float32)


def plot_modal_cloud_figure(
    result: ChannelModalCloudResult,
    analysis_settings: ModalCloudAnalysisSettings,
    plot_settings: ModalCloudPlotSettings,
    title: Optional[str] = None,
):
    """Scatter of (frequency, RT60) with an optional median curve (same content as reference :469-507)."""
    from .plotting import create_figure_and_axis, pyplot

    pyplot(optional=False)
    import matplotlib.ticker as mticker

    figure, axis = create_figure_and_axis(title=title)
    axis.set_xlabel("Frequency (Hz)")
    axis.set_ylabel(f"RT60 estimate (s) [{result.metric.upper()}]")
    nyquist = 0.5 * float(result.sample_rate_hz)
    f_lo = float(np.clip(analysis_settings.f_min_hz, 1.0, nyquist))
    f_hi = float(np.clip(analysis_settings.f_max_hz, f_lo, nyquist))
    axis.set_xscale("log")
    axis.set_xlim(f_lo, f_hi)
    ticks = [t for t in (20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000) if f_lo <= t <= f_hi]
    if ticks:
        axis.set_xticks(ticks)
    axis.xaxis.set_major_formatter(mticker.FuncFormatter(lambda x, pos: f"{int(x / 1000)}k" if x >= 1000 else f"{int(x)}"))
    axis.xaxis.set_minor_formatter(mticker.NullFormatter())
    axis.grid(True, which="both", linestyle=":", linewidth=0.5)
    if not result.points:
        axis.text(0.5, 0.5, "No valid points (insufficient decay range).", transform=axis.transAxes, ha="center")
        return figure
    f = np.array([p.centre_hz for p in result.points], dtype=np.float32)
    y = np.array([p.rt60_seconds for p in result.points], dtype=np.float32)
    axis.scatter(f, y, s=12, alpha=0.85, label=f"{result.channel_name} ({len(result.points)} pts)")
    if plot_settings.show_median_curve:
        med = _median_curve(result.points, plot_settings.median_octave_window)
        if med is not None:
            axis.plot(med[0], med[1], alpha=0.9, linestyle="-", label=f"{result.channel_name} median")
    if plot_settings.ylim_seconds is not None:
        axis.set_ylim(plot_settings.ylim_seconds[0], plot_settings.ylim_seconds[1])
    axis.legend(loc="best")
    return figure


def plot_modal_cloud_from_wav_file(
    input_wav_file_path: str | Path,
    analysis_settings: Optional[ModalCloudAnalysisSettings] = None,
    plot_settings: Optional[ModalCloudPlotSettings] = None,
    output_basename: Optional[str | Path] = None,
    show_interactive: bool = True,
) -> List[ChannelModalCloudResult]:
    """Analyse, then one `<basename>_modalcloud_<CH>.png` per channel when matplotlib is available (:510-551)."""
    from .plotting import finalize_and_show_or_save, pyplot

    analysis_settings = analysis_settings or ModalCloudAnalysisSettings()
    plot_settings = plot_settings or ModalCloudPlotSettings()
    results = analyse_modal_cloud_from_wav_file(input_wav_file_path, analysis_settings)
    if (output_basename is not None or show_interactive) and pyplot() is not None:
        for r in results:
            figure = plot_modal_cloud_figure(r, analysis_settings, plot_settings,
                                             title=f"Modal cloud — {input_wav_file_path} — {r.channel_name}")
            target = None
            if output_basename is not None:
                base = Path(output_basename)
                target = base.with_name(f"{base.stem}_modalcloud_{r.channel_name}.png").with_suffix(".png")
            finalize_and_show_or_save(figure, output_path=target, show_interactive=show_interactive)
    return results


def summarise_modal_cloud_results_text(results: List[ChannelModalCloudResult]) -> str:
    """Per channel: header line, then median / p90 / max of the RT60 points (reference :559-567)."""
    lines: List[str] = []
    for r in results:
        dur = float(r.analysis_length_samples) / float(r.sample_rate_hz)
        lines.append(f"[{r.channel_name}] metric={r.metric} start_sample={r.analysis_start_sample_index} "
                     f"dur={dur:.3f}s points={len(r.points)}")
        if r.points:
            rt = np.array([p.rt60_seconds for p in r.points], dtype=np.float64)
            lines.append(f"  rt60: median={np.median(rt):.3f}s  p90={np.percentile(rt, 90):.3f}s  max={np.max(rt):.3f}s")
    return "\n".join(lines)
