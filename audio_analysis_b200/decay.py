"""Decay analysis: Schroeder EDC, EDT / T20 / T30 line fits, RT60 (B200 path).

Drop-in mirror of the reference's `analyse/decay.py` API: the same settings and
result dataclasses, function names, argument meaning, error messages and text
summary.  The arithmetic (reference decay.py:115-260) runs in the fused CUDA
kernel behind `aa_decay_batch`; channels are analysed as a batch, so the
per-channel functions are batch-size-1 calls of the same kernels.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from . import _native as nat
from .io import get_analysis_channels, load_wav_file

# --------------------------------------------------------------------------------------
# settings and result records (field-for-field the reference's, decay.py:44-107)
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class DecayAnalysisSettings:
    use_mono_downmix_for_stereo: bool = False
    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    edc_floor_db: float = -120.0
    edc_epsilon: float = 1e-20
    fit_lower_limit_db: float = -80.0
    t20_range_db: Tuple[float, float] = (-5.0, -25.0)
    t30_range_db: Tuple[float, float] = (-5.0, -35.0)
    compute_edt: bool = False
    edt_range_db: Tuple[float, float] = (0.0, -10.0)
    edc_smoothing_window_samples: int = 0


@dataclass(frozen=True)
class LinearDecayFit:
    name: str
    range_db: Tuple[float, float]
    start_time_seconds: float
    end_time_seconds: float
    slope_db_per_second: float
    intercept_db: float
    r_squared: float
    rt60_seconds: float


@dataclass(frozen=True)
class ChannelDecayAnalysis:
    channel_name: str
    sample_rate_hz: int
    analysis_start_sample_index: int
    time_seconds: np.ndarray
    edc_db: np.ndarray
    early_decay_10db_time_seconds: Optional[float]
    fits: Dict[str, LinearDecayFit]


@dataclass(frozen=True)
class DecayPlotSettings:
    show_fit_lines: bool = True
    secondary_channel_alpha: float = 0.7
    ylim_db: Tuple[float, float] = (-120.0, 5.0)


_TOO_SHORT = "Not enough samples after trimming/ignoring to compute EDC."


# --------------------------------------------------------------------------------------
# settings -> kernel parameters, tables -> records
# --------------------------------------------------------------------------------------


def fit_plan(settings: DecayAnalysisSettings) -> List[Tuple[str, Tuple[float, float]]]:
    """Fits in the reference's evaluation order (decay.py:290-319): [EDT,] T20, T30."""
    plan: List[Tuple[str, Tuple[float, float]]] = []
    if settings.compute_edt:
        plan.append(("EDT", tuple(settings.edt_range_db)))
    plan.append(("T20", tuple(settings.t20_range_db)))
    plan.append(("T30", tuple(settings.t30_range_db)))
    return plan


def native_params(settings: DecayAnalysisSettings, sample_rate_hz: int,
                  fit_ranges: Sequence[Tuple[float, float]], want_early10: bool = True) -> nat.DecayParams:
    return engine.decay_params(
        sample_rate_hz=sample_rate_hz,
        trim_to_peak=bool(settings.trim_to_peak),
        ignore_leading_seconds=float(settings.ignore_leading_seconds),
        edc_floor_db=float(settings.edc_floor_db),
        edc_epsilon=float(settings.edc_epsilon),
        fit_lower_limit_db=float(settings.fit_lower_limit_db),
        fit_ranges=fit_ranges,
        want_early10=want_early10,
    )


def time_axis(count: int, sample_rate_hz: int) -> np.ndarray:
    """time_seconds of the reference (decay.py:169): float32 arange / sample rate."""
    return (np.arange(int(count), dtype=np.float32) / float(sample_rate_hz)).astype(np.float32)


def fit_from_row(name: str, range_db: Tuple[float, float], row: np.ndarray) -> Optional[LinearDecayFit]:
    """One [8] row of the kernel's fit table -> LinearDecayFit, or None when the fit is unavailable."""
    if row[7] != 1.0:
        return None
    return LinearDecayFit(
        name=name,
        range_db=(float(range_db[0]), float(range_db[1])),
        start_time_seconds=float(row[0]),
        end_time_seconds=float(row[1]),
        slope_db_per_second=float(row[2]),
        intercept_db=float(row[3]),
        r_squared=float(row[4]),
        rt60_seconds=float(row[5]),
    )


def run_decay(rows: engine.Rows, sample_rate_hz: int, settings: DecayAnalysisSettings, want_edc: bool = True,
              fit_ranges: Optional[Sequence[Tuple[float, float]]] = None, edc_out=None) -> engine.DecayOut:
    """Launch the decay kernels for a device batch (fused path, or the smoothing path of decay.py:161-164)."""
    ranges = [r for _, r in fit_plan(settings)] if fit_ranges is None else list(fit_ranges)
    params = native_params(settings, sample_rate_hz, ranges)
    window = int(settings.edc_smoothing_window_samples or 0)
    if window > 1:
        return engine.decay_smoothed(rows, params, window)
    return engine.decay(rows, params, want_edc=want_edc, edc_out=edc_out)


def analyse_decay_batch(
    channels,
    sample_rate_hz: int,
    channel_names: Optional[Sequence[str]] = None,
    settings: Optional[DecayAnalysisSettings] = None,
    device=None,
    return_curves: bool = True,
) -> List[ChannelDecayAnalysis]:
    """Analyse many channels in one launch.  `channels`: [B, N] array / tensor or a list of 1-D arrays.

    With return_curves=False the EDC stays on the device (time_seconds / edc_db are empty
    arrays); metrics are unaffected.
    """
    if settings is None:
        settings = DecayAnalysisSettings()
    rows = channels if isinstance(channels, engine.Rows) else engine.make_rows(channels, device)
    out = run_decay(rows, sample_rate_hz, settings, want_edc=return_curves)
    names = list(channel_names) if channel_names is not None else [f"ch{i}" for i in range(rows.n_rows)]
    return decay_results(out, rows, names, settings, sample_rate_hz, return_curves)


def decay_results(out: engine.DecayOut, rows: engine.Rows, names: Sequence[str], settings: DecayAnalysisSettings,
                  sample_rate_hz: int, return_curves: bool = True,
                  strict: bool = True) -> List[Optional[ChannelDecayAnalysis]]:
    """Result records from the device tables.  strict=False: a row the reference would raise on (fewer than 4 samples
    after trimming, decay.py:146) yields None instead of an exception (batched reports isolate such files)."""
    plan = fit_plan(settings)
    status = out.status.cpu().numpy()
    if strict and np.any(status == nat.AA_ROW_TOO_SHORT):
        raise ValueError(_TOO_SHORT)
    start = out.start.cpu().numpy()
    fits = out.fits.cpu().numpy()
    early = out.early10.cpu().numpy()
    curves = out.edc_db.cpu().numpy() if (return_curves and out.edc_db is not None) else None
    results: List[Optional[ChannelDecayAnalysis]] = []
    for b in range(rows.n_rows):
        if status[b] == nat.AA_ROW_TOO_SHORT:
            results.append(None)
            continue
        s, n = int(start[b]), int(rows.host_lengths[b])
        if curves is not None:
            edc = np.array(curves[b, s:n], dtype=np.float32)
            t = time_axis(n - s, sample_rate_hz)
        else:
            edc = np.zeros((0,), dtype=np.float32)
            t = np.zeros((0,), dtype=np.float32)
        found: Dict[str, LinearDecayFit] = {}
        for r, (name, rng) in enumerate(plan):
            fit = fit_from_row(name, rng, fits[b, r])
            if fit is not None:
                found[name] = fit
        e10 = None if np.isnan(early[b]) else float(early[b])
        results.append(ChannelDecayAnalysis(
            channel_name=str(names[b]),
            sample_rate_hz=int(sample_rate_hz),
            analysis_start_sample_index=s,
            time_seconds=t,
            edc_db=edc,
            early_decay_10db_time_seconds=e10,
            fits=found,
        ))
    return results


# --------------------------------------------------------------------------------------
# reference-named primitives
# --------------------------------------------------------------------------------------


def compute_schroeder_edc_db(
    samples: np.ndarray,
    sample_rate_hz: int,
    settings: DecayAnalysisSettings,
) -> Tuple[np.ndarray, np.ndarray, int]:
    """(time_seconds, edc_db, analysis_start_sample_index) - reference decay.py:115-170."""
    samples = np.asarray(samples)
    if samples.ndim != 1:
        raise ValueError("compute_schroeder_edc_db expects a 1D mono array.")
    rows = engine.make_rows(samples.astype(np.float32, copy=False)[None, :])
    out = run_decay(rows, sample_rate_hz, settings, want_edc=True, fit_ranges=[])
    if int(out.status.cpu()[0]) == nat.AA_ROW_TOO_SHORT:
        raise ValueError(_TOO_SHORT)
    start = int(out.start.cpu()[0])
    edc = out.edc_db[0, start:samples.size].cpu().numpy().astype(np.float32)
    return time_axis(edc.size, sample_rate_hz), edc, start


def _sample_rate_of(time_seconds: np.ndarray) -> int:
    t = np.asarray(time_seconds)
    if t.ndim != 1 or t.size < 2:
        raise ValueError("time_seconds must be a 1-D array with at least two samples")
    sr = int(round(1.0 / float(t[1] - t[0])))
    if sr <= 0 or not np.array_equal(time_axis(t.size, sr), t.astype(np.float32)):
        raise NotImplementedError(
            "fit_decay_slope_over_db_range on the device needs the uniform time axis arange(n)/sample_rate "
            "that compute_schroeder_edc_db returns"
        )
    return sr


def fit_decay_slope_over_db_range(
    time_seconds: np.ndarray,
    edc_db: np.ndarray,
    range_db: Tuple[float, float],
    fit_lower_limit_db: float,
    fit_name: str,
) -> Optional[LinearDecayFit]:
    """Least-squares line over the part of the curve between two dB levels (reference decay.py:202-260)."""
    hi, lo = float(range_db[0]), float(range_db[1])
    if lo > hi:
        raise ValueError("range_db should be (higher_db, lower_db), e.g. (-5, -25).")
    sr = _sample_rate_of(time_seconds)
    params = engine.decay_params(sr, False, 0.0, -np.inf, 0.0, float(fit_lower_limit_db), [(hi, lo)], False)
    curve = engine.make_rows(np.asarray(edc_db, dtype=np.float32)[None, :])
    fits, _ = engine.curve_fits(curve.x, None, curve.n, None, params)
    return fit_from_row(fit_name, (hi, lo), fits.cpu().numpy()[0, 0])


def analyse_decay_for_channel(
    samples: np.ndarray,
    sample_rate_hz: int,
    channel_name: str,
    settings: DecayAnalysisSettings,
) -> ChannelDecayAnalysis:
    """EDC + early-decay time + EDT/T20/T30 fits for one channel (reference decay.py:268-329)."""
    samples = np.asarray(samples)
    if samples.ndim != 1:
        raise ValueError("compute_schroeder_edc_db expects a 1D mono array.")
    return analyse_decay_batch(samples.astype(np.float32, copy=False)[None, :], sample_rate_hz, [channel_name],
                               settings)[0]


def analyse_decay_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[DecayAnalysisSettings] = None,
) -> List[ChannelDecayAnalysis]:
    """Load a WAV and analyse its channels in one batch (reference decay.py:333-365)."""
    if settings is None:
        settings = DecayAnalysisSettings()
    audio = load_wav_file(input_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    chans = get_analysis_channels(audio, use_mono_downmix_for_stereo=settings.use_mono_downmix_for_stereo)
    data = np.stack([c for _, c in chans], axis=0)
    return analyse_decay_batch(data, audio.sample_rate_hz, [n for n, _ in chans], settings)


# --------------------------------------------------------------------------------------
# figures (optional) and text
# --------------------------------------------------------------------------------------


def plot_decay_figure(
    channel_analyses: List[ChannelDecayAnalysis],
    analysis_settings: DecayAnalysisSettings,
    plot_settings: DecayPlotSettings,
    title: Optional[str] = None,
):
    """EDC curves with dashed fit lines and range guides (same content as reference decay.py:373-435)."""
    from .plotting import create_figure_and_axis, label_decibel_axis, label_time_axis_seconds

    figure, axis = create_figure_and_axis(title=title)
    label_time_axis_seconds(axis)
    label_decibel_axis(axis)
    axis.set_ylim(*plot_settings.ylim_db)
    for position, item in enumerate(channel_analyses):
        alpha = 1.0 if position == 0 else float(plot_settings.secondary_channel_alpha)
        axis.plot(item.time_seconds, item.edc_db, alpha=alpha, label=None)
        if not plot_settings.show_fit_lines:
            continue
        for key in ("EDT", "T20", "T30"):
            fit = item.fits.get(key)
            if fit is None:
                continue
            ends = np.array([fit.start_time_seconds, fit.end_time_seconds], dtype=np.float32)
            levels = (fit.slope_db_per_second * ends + fit.intercept_db).astype(np.float32)
            text = f"{fit.name} {item.channel_name}  {fit.rt60_seconds:.2f}s"
            if key == "EDT":
                e10 = item.early_decay_10db_time_seconds
                text += "  Δ10dB=NA" if e10 is None else f"  Δ10dB={e10:.3f}s"
            axis.plot(ends, levels, alpha=alpha, linestyle="--", label=text)
    for level in (analysis_settings.t20_range_db[0], analysis_settings.t20_range_db[1],
                  analysis_settings.t30_range_db[1], analysis_settings.fit_lower_limit_db):
        axis.axhline(float(level), linestyle=":", linewidth=1.0)
    axis.grid(True, which="both", linestyle=":", linewidth=0.5)
    axis.legend(loc="best")
    return figure


def plot_decay_from_wav_file(
    input_wav_file_path: str | Path,
    analysis_settings: Optional[DecayAnalysisSettings] = None,
    plot_settings: Optional[DecayPlotSettings] = None,
    output_basename: Optional[str | Path] = None,
    show_interactive: bool = True,
) -> List[ChannelDecayAnalysis]:
    """Analyse, then draw `<basename>_decay.png` when matplotlib is available (reference decay.py:451-494)."""
    from .plotting import finalize_and_show_or_save, pyplot

    analysis_settings = analysis_settings or DecayAnalysisSettings()
    plot_settings = plot_settings or DecayPlotSettings()
    results = analyse_decay_from_wav_file(input_wav_file_path, analysis_settings)
    if (output_basename is not None or show_interactive) and pyplot() is not None:
        figure = plot_decay_figure(results, analysis_settings, plot_settings, title=f"Decay (EDC) — {input_wav_file_path}")
        target = None
        if output_basename is not None:
            base = Path(output_basename)
            target = base.with_name(f"{base.stem}_decay.png").with_suffix(".png")
        finalize_and_show_or_save(figure, output_path=target, show_interactive=show_interactive)
    return results


def summarise_decay_results_text(channel_analyses: List[ChannelDecayAnalysis]) -> str:
    """Deterministic multi-line summary, byte-for-byte the reference's format (decay.py:502-542)."""
    out: List[str] = []
    for item in channel_analyses:
        out.append(f"[{item.channel_name}] analysis_start_sample_index={item.analysis_start_sample_index}")
        e10 = item.early_decay_10db_time_seconds
        out.append("  early_0_to_-10_time=NA" if e10 is None else f"  early_0_to_-10_time={e10:.4f}s")
        if not item.fits:
            out += ["  fits=NA", ""]
            continue
        for key in ("EDT", "T20", "T30"):
            fit = item.fits.get(key)
            if fit is None:
                out.append(f"  {key}: NA")
            else:
                out.append(
                    f"  {fit.name}: range=[{fit.range_db[0]:.1f},{fit.range_db[1]:.1f}]dB "
                    f"time=[{fit.start_time_seconds:.4f},{fit.end_time_seconds:.4f}]s "
                    f"slope={fit.slope_db_per_second:.6f}dB/s r2={fit.r_squared:.6f} rt60={fit.rt60_seconds:.4f}s"
                )
        out.append("")
    return "\n".join(out).rstrip() + "\n"
