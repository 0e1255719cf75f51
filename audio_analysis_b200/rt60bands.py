"""Band-limited RT60 (three-band / octave / third-octave) on the B200 path.

Drop-in mirror of the reference's `analyse/rt60bands.py` API.  The band table is
host arithmetic (a handful of doubles, reference rt60bands.py:183-264).  The
per-band work - raised-cosine mask, circular FFT filter over the whole buffer,
Schroeder EDC and line fits (reference rt60bands.py:116-175, :272-321, :362-406) -
runs behind `aa_rt60bands_batch` for all channels and bands at once, with ONE
forward transform per channel.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, replace
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from . import _native as nat
from .decay import DecayAnalysisSettings, native_params
from .io import get_analysis_channels, load_wav_file


@dataclass(frozen=True)
class Rt60BandsAnalysisSettings:
    band_mode: str = "three"  # "three" | "octave" | "third"
    low_upper_hz: float = 250.0
    mid_center_hz: float = 1000.0
    mid_width_octaves: float = 2.0
    high_lower_hz: float = 4000.0
    f_min_hz: float = 31.5
    f_max_hz: float = 16000.0
    transition_width_octaves: float = (1.0 / 6.0)
    include_t20: bool = False
    include_edt: bool = False
    decay_settings: DecayAnalysisSettings = DecayAnalysisSettings()
    # Extension of this build (not a reference field; defaults reproduce the reference): the band filter.
    #   "fft_mask"    - the reference's circular FFT-domain raised-cosine mask (rt60bands.py:116-175)
    #   "butterworth" - Butterworth sections of `butterworth_order`, zero-phase forward-backward filtering
    #                   (scipy.signal.sosfiltfilt semantics; csrc/sos.cu).  Results differ from the reference's by
    #                   0.04 - 11 % depending on the band: parity with the reference is NOT claimed.
    filterbank: str = "fft_mask"
    butterworth_order: int = 4


@dataclass(frozen=True)
class Rt60BandsPlotSettings:
    ylim_seconds: Optional[Tuple[float, float]] = None
    secondary_channel_alpha: float = 0.7
    legend_values: bool = True


@dataclass(frozen=True)
class BandDefinition:
    name: str
    centre_hz: float
    kind: str  # "lowpass" | "bandpass" | "highpass"
    low_edge_hz: Optional[float] = None
    high_edge_hz: Optional[float] = None


@dataclass(frozen=True)
class Rt60BandMetrics:
    rt60_t30_seconds: Optional[float]
    rt60_t20_seconds: Optional[float]
    edt_seconds: Optional[float]


@dataclass(frozen=True)
class Rt60BandsChannelResult:
    channel_name: str
    sample_rate_hz: int
    band_definitions: List[BandDefinition]
    band_metrics_by_name: Dict[str, Rt60BandMetrics]


_KIND_CODE = {"lowpass": nat.AA_BAND_LOWPASS, "bandpass": nat.AA_BAND_BANDPASS, "highpass": nat.AA_BAND_HIGHPASS}


# --------------------------------------------------------------------------------------
# band tables (host; same double-precision expressions as the reference so the table is
# reproduced bit for bit - e.g. the absent 16 kHz third-octave band)
# --------------------------------------------------------------------------------------


def _clamp(value: float, lo: float, hi: float) -> float:
    return float(min(max(float(value), lo), hi))


def _octave_factor(octaves: float) -> float:
    return float(2.0 ** float(octaves))


def _build_three_band_definitions(settings: Rt60BandsAnalysisSettings, sample_rate_hz: int) -> List[BandDefinition]:
    """Low (lowpass) / Mid (bandpass) / High (highpass) - reference rt60bands.py:183-205."""
    nyquist = 0.5 * float(sample_rate_hz)
    low_top = _clamp(settings.low_upper_hz, 20.0, nyquist)
    mid_c = _clamp(settings.mid_center_hz, 20.0, nyquist)
    spread = _octave_factor(0.5 * float(max(0.1, settings.mid_width_octaves)))
    high_bottom = _clamp(settings.high_lower_hz, 20.0, nyquist)
    return [
        BandDefinition("Low", float(math.sqrt(20.0 * low_top)), "lowpass", high_edge_hz=low_top),
        BandDefinition("Mid", mid_c, "bandpass", low_edge_hz=_clamp(mid_c / spread, 20.0, nyquist),
                       high_edge_hz=_clamp(mid_c * spread, 20.0, nyquist)),
        BandDefinition("High", float(math.sqrt(max(20.0, high_bottom) * nyquist)), "highpass",
                       low_edge_hz=high_bottom),
    ]


def _build_fractional_octave_band_definitions(settings: Rt60BandsAnalysisSettings, sample_rate_hz: int,
                                              bands_per_octave: int) -> List[BandDefinition]:
    """Centres 1000 * 2^(k/n) inside [f_min, f_max], edges fc * 2^(+-1/2n) - reference rt60bands.py:208-253."""
    nyquist = 0.5 * float(sample_rate_hz)
    f_lo = float(max(20.0, min(settings.f_min_hz, nyquist)))
    f_hi = float(max(f_lo, min(settings.f_max_hz, nyquist)))
    n = float(bands_per_octave)
    step = 2.0 ** (1.0 / n)
    half_band = 2.0 ** (1.0 / (2.0 * n))
    k_first = int(np.floor(np.log(f_lo / 1000.0) / np.log(step)))
    k_last = int(np.ceil(np.log(f_hi / 1000.0) / np.log(step)))
    table: List[BandDefinition] = []
    for k in range(k_first, k_last + 1):
        centre = 1000.0 * (step ** float(k))
        if centre < f_lo or centre > f_hi:
            continue
        lower = _clamp(centre / half_band, 20.0, nyquist)
        upper = _clamp(centre * half_band, 20.0, nyquist)
        if upper <= lower:
            continue
        table.append(BandDefinition(f"{int(round(centre))}Hz", float(centre), "bandpass", lower, upper))
    table.sort(key=lambda band: band.centre_hz)
    return table


def _build_band_definitions(settings: Rt60BandsAnalysisSettings, sample_rate_hz: int) -> List[BandDefinition]:
    mode = str(settings.band_mode).lower()
    if mode == "three":
        return _build_three_band_definitions(settings, sample_rate_hz)
    if mode == "octave":
        return _build_fractional_octave_band_definitions(settings, sample_rate_hz, 1)
    if mode == "third":
        return _build_fractional_octave_band_definitions(settings, sample_rate_hz, 3)
    raise ValueError(f"Unknown band_mode: {settings.band_mode}")


# --------------------------------------------------------------------------------------
# batched analysis
# --------------------------------------------------------------------------------------


def band_fit_plan(settings: Rt60BandsAnalysisSettings) -> List[Tuple[str, Tuple[float, float]]]:
    """Kernel fit order for the band table: T30 always, then T20 / EDT when requested (rt60bands.py:285-315)."""
    ds = settings.decay_settings
    plan = [("T30", tuple(ds.t30_range_db))]
    if settings.include_t20:
        plan.append(("T20", tuple(ds.t20_range_db)))
    if settings.include_edt:
        plan.append(("EDT", tuple(ds.edt_range_db)))
    return plan


def band_triples(bands: Sequence[BandDefinition]):
    out = []
    for b in bands:
        if b.kind not in _KIND_CODE:
            raise ValueError(f"Unknown band kind: {b.kind}")
        out.append((_KIND_CODE[b.kind], b.low_edge_hz, b.high_edge_hz))
    return out


def decimation_classes(bands: Sequence[BandDefinition], n: int, sample_rate_hz: int, transition_width_octaves: float,
                       device=None, settings: Optional[Rt60BandsAnalysisSettings] = None) -> List[int]:
    """Decimation factor the device path uses for each band of a length-`n` row in a metrics-only call
    (1 = full rate), as classified by the library (`aa_bands_classify`, csrc/bands.cu) under the current
    `aa_bands_set_decimation` limit.  Used for byte accounting (bench.py) and documentation."""
    import torch

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    params = None
    if settings is not None:
        params = native_params(settings.decay_settings, sample_rate_hz, [r for _, r in band_fit_plan(settings)],
                               want_early10=False)
    return engine.bands_classify(n, dev, band_triples(bands), float(transition_width_octaves), int(sample_rate_hz), params)


def run_rt60_bands(rows: engine.Rows, sample_rate_hz: int, settings: Rt60BandsAnalysisSettings,
                   start_idx=None, bands: Optional[Sequence[BandDefinition]] = None, want_band_signals: bool = False,
                   options=None, workspace_bytes: Optional[int] = None):
    """Device-level call: returns (metrics tensor [B, nb, 3], bands, plan, start_idx, band signals or None).
    `options`: engine.bands_options(...) - decimation / certification of the band stage (None = defaults)."""
    if rows.n < 8:  # reference rt60bands.py:346-347
        raise ValueError("Not enough samples for rt60bands analysis.")
    bands = list(bands) if bands is not None else _build_band_definitions(settings, sample_rate_hz)
    plan = band_fit_plan(settings)
    ds = settings.decay_settings
    params = native_params(ds, sample_rate_hz, [r for _, r in plan], want_early10=False)
    params.smoothing_window = int(ds.edc_smoothing_window_samples or 0)
    if start_idx is None:
        # one time-zero for every band, from the full-band signal (reference rt60bands.py:333-342)
        import torch

        if ds.trim_to_peak:
            start_idx = engine.peak_index(rows)
        else:
            start_idx = torch.zeros(rows.n_rows, dtype=torch.int32, device=rows.x.device)
        ign = int(params.ignore_leading_samples)
        if ign > 0:
            start_idx = torch.clamp(start_idx + min(ign, rows.n), max=rows.n).to(torch.int32)
    mode = str(getattr(settings, "filterbank", "fft_mask")).lower()
    if mode == "butterworth":
        from . import butterworth

        if int(params.smoothing_window) > 1:
            raise ValueError("edc_smoothing_window_samples is not available with filterbank='butterworth'")
        nyq = 0.5 * float(sample_rate_hz)
        triples = [(b.kind, None if b.low_edge_hz is None else float(np.clip(b.low_edge_hz, 1.0, nyq - 1.0)),
                    None if b.high_edge_hz is None else float(np.clip(b.high_edge_hz, 1.0, nyq - 1.0))) for b in bands]
        sos, zi, edges = butterworth.band_sections(triples, int(settings.butterworth_order), float(sample_rate_hz))
        metrics, sig = engine.rt60bands_sos(rows, start_idx, sos, zi, edges, params,
                                            want_band_signals=want_band_signals, workspace_bytes=workspace_bytes)
        return metrics, bands, plan, start_idx, sig
    if mode != "fft_mask":
        raise ValueError(f"Unknown filterbank: {settings.filterbank}")
    metrics, sig = engine.rt60bands(rows, start_idx, band_triples(bands), float(settings.transition_width_octaves),
                                    params, want_band_signals=want_band_signals, workspace_bytes=workspace_bytes,
                                    options=options)
    return metrics, bands, plan, start_idx, sig


def metrics_to_results(table: np.ndarray, bands: Sequence[BandDefinition], plan, names: Sequence[str],
                       sample_rate_hz: int) -> List[Rt60BandsChannelResult]:
    col = {name: i for i, (name, _) in enumerate(plan)}
    # one conversion to Python floats for the whole table (NaN -> None), then plain indexing
    vals = np.where(np.isnan(table), None, table.astype(object)).tolist()
    c30, c20, cedt = col.get("T30"), col.get("T20"), col.get("EDT")
    band_list = list(bands)
    results = []
    for b, name in enumerate(names):
        rows_b = vals[b]
        by_name = {
            band.name: Rt60BandMetrics(None if c30 is None else rows_b[k][c30], None if c20 is None else rows_b[k][c20],
                                       None if cedt is None else rows_b[k][cedt])
            for k, band in enumerate(band_list)
        }
        results.append(Rt60BandsChannelResult(str(name), int(sample_rate_hz), band_list, by_name))
    return results


def analyse_rt60_bands_batch(channels, sample_rate_hz: int, channel_names: Optional[Sequence[str]] = None,
                             settings: Optional[Rt60BandsAnalysisSettings] = None,
                             device=None) -> List[Rt60BandsChannelResult]:
    """All channels x all bands in one call.  `channels`: [B, N] (equal lengths) array / tensor / Rows."""
    settings = settings or Rt60BandsAnalysisSettings()
    rows = channels if isinstance(channels, engine.Rows) else engine.make_rows(channels, device)
    names = list(channel_names) if channel_names is not None else [f"ch{i}" for i in range(rows.n_rows)]
    if not rows.uniform:
        # group by length: the FFT plan is per length
        out: List[Optional[Rt60BandsChannelResult]] = [None] * rows.n_rows
        for length in sorted(set(int(v) for v in rows.host_lengths)):
            idx = [i for i, v in enumerate(rows.host_lengths) if int(v) == length]
            sub = engine.make_rows(rows.x[idx, :length].contiguous(), rows.x.device)
            for i, r in zip(idx, analyse_rt60_bands_batch(sub, sample_rate_hz, [names[i] for i in idx], settings)):
                out[i] = r
        return out  # type: ignore[return-value]
    metrics, bands, plan, _, _ = run_rt60_bands(rows, sample_rate_hz, settings)
    return metrics_to_results(metrics.cpu().numpy(), bands, plan, names, sample_rate_hz)


def analyse_rt60_bands_for_channel(
    samples: np.ndarray,
    sample_rate_hz: int,
    channel_name: str,
    settings: Rt60BandsAnalysisSettings,
) -> Rt60BandsChannelResult:
    """One channel, every band (reference rt60bands.py:324-413)."""
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("analyse_rt60_bands_for_channel expects a 1D mono array.")
    if samples.size < 8:
        raise ValueError("Not enough samples for rt60bands analysis.")
    return analyse_rt60_bands_batch(samples[None, :], sample_rate_hz, [channel_name], settings)[0]


def analyse_rt60_bands_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[Rt60BandsAnalysisSettings] = None,
) -> List[Rt60BandsChannelResult]:
    """Reference rt60bands.py:417-445; both channels of a stereo file share one batch."""
    settings = settings or Rt60BandsAnalysisSettings()
    audio = load_wav_file(input_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    chans = get_analysis_channels(audio, use_mono_downmix_for_stereo=settings.decay_settings.use_mono_downmix_for_stereo)
    data = np.stack([c for _, c in chans], axis=0)
    return analyse_rt60_bands_batch(data, audio.sample_rate_hz, [n for n, _ in chans], settings)


# --------------------------------------------------------------------------------------
# figure (optional) and text table
# --------------------------------------------------------------------------------------


def _metric_names(include_t20: bool, include_edt: bool) -> List[str]:
    return ["T30"] + (["T20"] if include_t20 else []) + (["EDT"] if include_edt else [])


def _metric_of(m: Optional[Rt60BandMetrics], key: str) -> Optional[float]:
    if m is None:
        return None
    return {"T30": m.rt60_t30_seconds, "T20": m.rt60_t20_seconds, "EDT": m.edt_seconds}[key]


def plot_rt60_bands_figure(
    channel_results: List[Rt60BandsChannelResult],
    settings: Rt60BandsAnalysisSettings,
    plot_settings: Rt60BandsPlotSettings,
    title: Optional[str] = None,
):
    """<= 6 bands: grouped bars; otherwise lines over log frequency (same content as reference :453-577)."""
    from .plotting import create_figure_and_axis

    if not channel_results:
        raise ValueError("No channel results to plot.")
    bands = channel_results[0].band_definitions
    labels = [b.name for b in bands]
    keys = _metric_names(settings.include_t20, settings.include_edt)
    figure, axis = create_figure_and_axis(title=title)
    axis.set_ylabel("RT60 (seconds)")
    series = []
    for ci, ch in enumerate(channel_results):
        alpha = 1.0 if ci == 0 else float(plot_settings.secondary_channel_alpha)
        for key in keys:
            vals = [_metric_of(ch.band_metrics_by_name.get(n), key) for n in labels]
            text = f"{key} {ch.channel_name}"
            if plot_settings.legend_values:
                text += "  " + "  ".join(f"{n}={'NA' if v is None else f'{v:.2f}s'}" for n, v in zip(labels, vals))
            series.append((np.array([np.nan if v is None else v for v in vals], dtype=np.float32), alpha, text, key))
    if len(bands) <= 6:
        axis.set_xlabel("Band")
        xs = np.arange(len(bands), dtype=np.float32)
        axis.set_xticks(xs)
        axis.set_xticklabels(labels)
        width = 0.8 / max(1, len(series))
        for si, (vals, alpha, text, _) in enumerate(series):
            axis.bar(xs + (si - len(series) / 2) * width + width / 2, vals, width=width, alpha=alpha, label=text)
        axis.grid(True, axis="y", linestyle=":", linewidth=0.5)
    else:
        axis.set_xlabel("Band centre frequency (Hz)")
        axis.set_xscale("log")
        axis.grid(True, which="both", linestyle=":", linewidth=0.5)
        centres = np.array([b.centre_hz for b in bands], dtype=np.float32)
        styles = {"T30": "-", "T20": "--", "EDT": ":"}
        for vals, alpha, text, key in series:
            axis.plot(centres, vals, linestyle=styles[key], marker="o", alpha=alpha, label=text)
    if plot_settings.ylim_seconds is not None:
        axis.set_ylim(*plot_settings.ylim_seconds)
    axis.legend(loc="best")
    return figure


def plot_rt60_bands_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[Rt60BandsAnalysisSettings] = None,
    plot_settings: Optional[Rt60BandsPlotSettings] = None,
    output_basename: Optional[str | Path] = None,
    show_interactive: bool = True,
) -> List[Rt60BandsChannelResult]:
    """Analyse, then draw `<basename>_rt60bands.png` when matplotlib is available (reference :580-624)."""
    from .plotting import finalize_and_show_or_save, pyplot

    settings = settings or Rt60BandsAnalysisSettings()
    plot_settings = plot_settings or Rt60BandsPlotSettings()
    if plot_settings.legend_values and str(settings.band_mode).lower() in ("octave", "third"):
        plot_settings = replace(plot_settings, legend_values=False)  # dense modes: names only (reference :592-597)
    results = analyse_rt60_bands_from_wav_file(input_wav_file_path, settings)
    if (output_basename is not None or show_interactive) and pyplot() is not None:
        figure = plot_rt60_bands_figure(results, settings, plot_settings, title=f"RT60 bands — {input_wav_file_path}")
        target = None
        if output_basename is not None:
            base = Path(output_basename)
            target = base.with_name(f"{base.stem}_rt60bands.png").with_suffix(".png")
        finalize_and_show_or_save(figure, output_path=target, show_interactive=show_interactive)
    return results


def summarise_rt60_bands_results_text(
    channel_results: List[Rt60BandsChannelResult],
    include_t20: bool,
    include_edt: bool,
) -> str:
    """`Band  T30_RT60(s) [...]` table per channel, 3 decimals, NA for None (reference :627-666)."""
    keys = _metric_names(include_t20, include_edt)
    out: List[str] = []
    for ch in channel_results:
        out.append(f"[{ch.channel_name}]")
        out.append("  ".join(["Band"] + [f"{k}_RT60(s)" for k in keys]))
        for band in ch.band_definitions:
            m = ch.band_metrics_by_name.get(band.name)
            cells = [band.name]
            for k in keys:
                v = _metric_of(m, k)
                cells.append("NA" if v is None else f"{float(v):.3f}")
            out.append("  ".join(cells))
        out.append("")
    return "\n".join(out)
