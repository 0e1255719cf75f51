"""Report writer for the accelerated blocks (decay, rt60bands, spectrogram, waterfall, modal cloud).

Mirror of the part of the reference's `analyse/report.py` that lies on the
accelerated path (report.py:252-287, :322-354, :371-386): same `ReportSettings`
fields, same Markdown section layout and order, same text summaries, same output
file names.  The other four report blocks (impulse-response plots, frequency
response, group delay, diffusion) are outside this build's scope (SURVEY.md
section 2); their flags are accepted and the report says explicitly that they were
not run, instead of silently dropping them.

`run_reports_batch` analyses many WAV files at once: every channel of every file
goes through one decay launch, one rt60bands call and one CSD launch.
"""

from __future__ import annotations

import dataclasses
from dataclasses import dataclass, replace
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from .decay import (ChannelDecayAnalysis, DecayAnalysisSettings, DecayPlotSettings, analyse_decay_batch,
                    plot_decay_figure, summarise_decay_results_text)
from .io import DEFAULT_EXPECTED_SAMPLE_RATE_HZ, LoadedAudio, get_analysis_channels, load_wav_file
from .modalcloud import (ChannelModalCloudResult, ModalCloudAnalysisSettings, ModalCloudPlotSettings,
                         analyse_modal_cloud_batch, plot_modal_cloud_figure, summarise_modal_cloud_results_text)
from .rt60bands import (Rt60BandsAnalysisSettings, Rt60BandsChannelResult, Rt60BandsPlotSettings,
                        analyse_rt60_bands_batch, plot_rt60_bands_figure, summarise_rt60_bands_results_text)
from .spectrogram import (ChannelSpectrogramResult, SpectrogramAnalysisSettings, SpectrogramPlotSettings,
                          analyse_spectrogram_batch, plot_spectrogram_figure, summarise_spectrogram_results_text)
from .waterfall import (ChannelWaterfallResult, WaterfallAnalysisSettings, WaterfallPlotSettings,
                        analyse_waterfall_batch, plot_waterfall_figure, summarise_waterfall_results_text)

OUT_OF_SCOPE_BLOCKS = (
    ("run_impulse_response_plots", "Impulse response"),
    ("run_frequency_response", "Frequency response"),
    ("run_group_delay", "Group delay"),
    ("run_diffusion", "Diffusion / echo density proxy"),
)


@dataclass(frozen=True)
class ReportSettings:
    common_use_mono_downmix_for_stereo: bool = False
    common_trim_to_peak: bool = True
    common_ignore_leading_seconds: float = 0.0

    run_impulse_response_plots: bool = True
    run_decay: bool = True
    run_rt60_bands: bool = True
    run_frequency_response: bool = True
    run_group_delay: bool = True
    run_spectrogram: bool = True
    run_waterfall: bool = True
    run_diffusion: bool = True
    run_modal_cloud: bool = True
    run_echo_density: bool = True

    expected_sample_rate_hz: int = DEFAULT_EXPECTED_SAMPLE_RATE_HZ

    ir_view_settings: Optional[Any] = None
    decay_analysis_settings: Optional[DecayAnalysisSettings] = None
    decay_plot_settings: Optional[DecayPlotSettings] = None
    rt60_bands_settings: Optional[Rt60BandsAnalysisSettings] = None
    rt60_bands_plot_settings: Optional[Rt60BandsPlotSettings] = None
    frequency_response_analysis_settings: Optional[Any] = None
    frequency_response_plot_settings: Optional[Any] = None
    group_delay_analysis_settings: Optional[Any] = None
    group_delay_plot_settings: Optional[Any] = None
    spectrogram_analysis_settings: Optional[SpectrogramAnalysisSettings] = None
    spectrogram_plot_settings: Optional[SpectrogramPlotSettings] = None
    waterfall_analysis_settings: Optional[WaterfallAnalysisSettings] = None
    waterfall_plot_settings: Optional[WaterfallPlotSettings] = None
    diffusion_analysis_settings: Optional[Any] = None
    modal_cloud_analysis_settings: Optional[ModalCloudAnalysisSettings] = None
    modal_cloud_plot_settings: Optional[ModalCloudPlotSettings] = None


@dataclass(frozen=True)
class ReportResults:
    input_wav_file_path: Path
    output_basename: Path
    summary_markdown_path: Path
    summary_markdown: str


def _apply_common_overrides(settings_obj: Any, report_settings: ReportSettings) -> Any:
    """Push the three common controls into any settings dataclass that has them (reference report.py:172-186)."""
    if settings_obj is None:
        return None
    names = {f.name for f in dataclasses.fields(settings_obj)}
    common = {
        "use_mono_downmix_for_stereo": report_settings.common_use_mono_downmix_for_stereo,
        "trim_to_peak": report_settings.common_trim_to_peak,
        "ignore_leading_seconds": report_settings.common_ignore_leading_seconds,
    }
    changes = {k: v for k, v in common.items() if k in names}
    return replace(settings_obj, **changes) if changes else settings_obj


def resolved_settings(settings: ReportSettings):
    """(decay, rt60bands, waterfall) analysis settings after the common overrides."""
    decay = _apply_common_overrides(settings.decay_analysis_settings or DecayAnalysisSettings(), settings)
    bands = _apply_common_overrides(settings.rt60_bands_settings or Rt60BandsAnalysisSettings(), settings)
    water = _apply_common_overrides(settings.waterfall_analysis_settings or WaterfallAnalysisSettings(), settings)
    return decay, bands, water


def _section(title: str) -> str:
    return f"\n## {title}\n\n"


def _code(text: str) -> str:
    text = text.strip()
    return f"```text\n{text}\n```\n" if text else "_(no output)_\n"


def _image(basename: Path, suffix: str, alt: str) -> str:
    return f"![{alt}]({basename.name}{suffix}.png)\n\n"


def _header(path: Path, audio_shape: Tuple[int, int], sample_rate: int, expected: int) -> str:
    n, ch = audio_shape
    duration = n / sample_rate if sample_rate > 0 else 0.0
    return (
        "# Offline Reverb Analysis Report\n\n"
        f"**Input WAV:** `{path}`  \n"
        f"**Sample rate:** {sample_rate} Hz (expected {expected} Hz)  \n"
        f"**Channels:** {ch}  \n"
        f"**Samples:** {n}  \n"
        f"**Duration:** {duration:.6f} s\n\n"
        "---\n"
    )


@dataclass
class _FileJob:
    path: Path
    basename: Path
    shape: Tuple[int, int]     # (samples, channels) of the file
    sample_rate_hz: int
    names: List[str]           # analysed channel names
    first_row: int = 0         # index of this file's first channel in its group's batch


def _per_channel_block(parts, have_plots, results, figure_fn, title, stem, alt, base, mono, summary) -> None:
    """One `<base>_<stem>_<CH>.png` per channel, the left / right image links and the text summary."""
    from .plotting import finalize_and_show_or_save

    if have_plots:
        for r in results:
            finalize_and_show_or_save(figure_fn(r), base.with_name(f"{base.stem}_{stem}_{r.channel_name}.png")
                                      .with_suffix(".png"), False)
    parts.append(_section(title))
    parts.append(_image(base, f"_{stem}_left", f"{alt} (left)"))
    if not mono:
        parts.append(_image(base, f"_{stem}_right", f"{alt} (right)"))
    parts.append(_code(summary))


def _render(job: _FileJob, settings: ReportSettings, decay_s, bands_s, water_s,
            decay_r: Optional[List[ChannelDecayAnalysis]], bands_r: Optional[List[Rt60BandsChannelResult]],
            water_r: Optional[List[ChannelWaterfallResult]], spec_s=None,
            spec_r: Optional[List[ChannelSpectrogramResult]] = None, modal_s=None,
            modal_r: Optional[List[ChannelModalCloudResult]] = None) -> ReportResults:
    from .plotting import finalize_and_show_or_save, pyplot

    have_plots = pyplot() is not None
    job.basename.parent.mkdir(parents=True, exist_ok=True)
    # header counts channels the way the reference does: mono files are shown upmixed to stereo (report.py:193-199)
    shape = (int(job.shape[0]), max(2, int(job.shape[1])))
    parts = [_header(job.path, shape, job.sample_rate_hz, settings.expected_sample_rate_hz)]
    base = job.basename
    if decay_r is not None:
        if have_plots:
            fig = plot_decay_figure(decay_r, decay_s, settings.decay_plot_settings or DecayPlotSettings(),
                                    title=f"Decay (EDC) — {job.path}")
            finalize_and_show_or_save(fig, base.with_name(f"{base.stem}_decay.png").with_suffix(".png"), False)
        parts += [_section("Decay / EDC"), _image(base, "_decay", "Decay analysis (T20/T30/RT60/EDT)"),
                  _code(summarise_decay_results_text(decay_r))]
    if bands_r is not None:
        if have_plots:
            ps = settings.rt60_bands_plot_settings or Rt60BandsPlotSettings()
            if ps.legend_values and str(bands_s.band_mode).lower() in ("octave", "third"):
                ps = replace(ps, legend_values=False)
            fig = plot_rt60_bands_figure(bands_r, bands_s, ps, title=f"RT60 bands — {job.path}")
            finalize_and_show_or_save(fig, base.with_name(f"{base.stem}_rt60bands.png").with_suffix(".png"), False)
        parts += [_section("RT60 by band"), _image(base, "_rt60bands", "RT60 by frequency band"),
                  _code(summarise_rt60_bands_results_text(bands_r, bool(bands_s.include_t20),
                                                          bool(bands_s.include_edt)))]
    if spec_r is not None:   # reference report.py:322-337
        ps = settings.spectrogram_plot_settings or SpectrogramPlotSettings()
        _per_channel_block(parts, have_plots, spec_r,
                           lambda r: plot_spectrogram_figure(r, spec_s, ps, title=f"Spectrogram — {job.path} — {r.channel_name}"),
                           "Spectrogram", "spectrogram", "Spectrogram", base,
                           settings.common_use_mono_downmix_for_stereo, summarise_spectrogram_results_text(spec_r))
    if water_r is not None:
        if have_plots:
            for r in water_r:
                fig = plot_waterfall_figure(r, water_s, settings.waterfall_plot_settings or WaterfallPlotSettings(),
                                            title=f"Waterfall — {job.path} — {r.channel_name}")
                finalize_and_show_or_save(
                    fig, base.with_name(f"{base.stem}_waterfall_{r.channel_name}.png").with_suffix(".png"), False)
        parts.append(_section("Waterfall"))
        parts.append(_image(base, "_waterfall_left", "Waterfall plot (left)"))
        if not settings.common_use_mono_downmix_for_stereo:
            parts.append(_image(base, "_waterfall_right", "Waterfall plot (right)"))
        parts.append(_code(summarise_waterfall_results_text(water_r)))
    if modal_r is not None:   # reference report.py:371-386
        ps = settings.modal_cloud_plot_settings or ModalCloudPlotSettings()
        _per_channel_block(parts, have_plots, modal_r,
                           lambda r: plot_modal_cloud_figure(r, modal_s, ps, title=f"Modal cloud — {job.path} — {r.channel_name}"),
                           "Modal cloud", "modalcloud", "Modal cloud", base,
                           settings.common_use_mono_downmix_for_stereo, summarise_modal_cloud_results_text(modal_r))
    skipped = [title for flag, title in OUT_OF_SCOPE_BLOCKS if getattr(settings, flag)]
    if skipped:
        parts.append(_section("Blocks not run"))
        parts.append("The B200 build accelerates the decay, RT60-by-band, spectrogram, waterfall and modal-cloud "
                     "blocks only. Not run: " + ", ".join(skipped) + ".\n")
    text = "".join(parts).rstrip() + "\n"
    md_path = Path(f"{job.basename}_report.md")
    md_path.parent.mkdir(parents=True, exist_ok=True)
    md_path.write_text(text, encoding="utf-8")
    return ReportResults(job.path, job.basename, md_path, text)


# rows per launch of the dense-STFT blocks (spectrogram n_fft 4096: 2.3 MB per second of audio and row, modal cloud
# n_fft 8192 the same again): bounds the [rows, frames, bins] tables to a few GB whatever the bundle size
DENSE_BLOCK_ROWS = 64


def _channel_names(channels: int, mono_downmix: bool) -> List[str]:
    if channels == 1 or mono_downmix:
        return ["mono"]
    return ["left", "right"]


def _analyse_group(rows: engine.Rows, sr: int, names: List[str], settings: ReportSettings, decay_s, bands_s, water_s,
                   spec_s, modal_s):
    """Every enabled block for one batch of equal-rate channels.  Returns per-block lists with one entry per row; an
    entry is None (or a ValueError for the dense blocks) where the reference would have raised for that channel."""
    from .decay import decay_results
    from .pipeline import run_hot_path
    from .rt60bands import metrics_to_results
    from .waterfall import waterfall_results

    out = {"decay": None, "bands": None, "water": None, "spec": None, "modal": None}
    smoothing = int(decay_s.edc_smoothing_window_samples or 0) > 1
    do_bands = bool(settings.run_rt60_bands) and rows.uniform
    # the curve-writing mode of the decay kernel always runs (same arithmetic, hence the same printed digits, as the
    # per-file functions); whether the curves then LEAVE the device is decided below
    hot = run_hot_path(rows, sr, decay_s, bands_s, water_s, want_edc=bool(settings.run_decay),
                       do_bands=do_bands, do_csd=bool(settings.run_waterfall), validate=False)
    if settings.run_decay:
        # the EDC curves are plot data: they leave the device only when there is a backend to draw them
        from .plotting import pyplot

        out["decay"] = decay_results(hot.decay, rows, names, decay_s, sr, return_curves=pyplot() is not None,
                                     strict=False)
    if settings.run_rt60_bands:
        if do_bands:
            out["bands"] = metrics_to_results(hot.band_metrics.cpu().numpy(), hot.bands, hot.band_fit_plan, names, sr)
        else:   # ragged group: the band stage plans one FFT per length
            out["bands"] = analyse_rt60_bands_batch(rows, sr, names, bands_s)
    if settings.run_waterfall:
        out["water"] = waterfall_results(hot.csd, hot.csd_plan, names, sr, strict=False)
    del hot
    peaks = None
    for key, enabled, fn, s in (("spec", settings.run_spectrogram, analyse_spectrogram_batch, spec_s),
                                ("modal", settings.run_modal_cloud, analyse_modal_cloud_batch, modal_s)):
        if not enabled:
            continue
        res: List[Any] = []
        for r0 in range(0, rows.n_rows, DENSE_BLOCK_ROWS):
            r1 = min(rows.n_rows, r0 + DENSE_BLOCK_ROWS)
            part = engine.slice_rows(rows, r0, r1)
            try:
                res += fn(part, sr, names[r0:r1], s)
            except ValueError:
                # isolate the channel(s) the reference would have raised on
                for r in range(r0, r1):
                    try:
                        res += fn(engine.slice_rows(rows, r, r + 1), sr, names[r:r + 1], s)
                    except ValueError as exc:
                        res.append(exc)
        out[key] = res
    return out


def run_reports_batch(
    input_wav_file_paths: Sequence[str | Path],
    output_basenames: Sequence[str | Path],
    settings: Optional[ReportSettings] = None,
    loaded: Optional[Sequence[LoadedAudio]] = None,
) -> List[ReportResults]:
    """Reports for many WAV files with every channel analysed in shared launches.

    PCM16 files of one common shape (a bundle's taps, recorder.hpp:55-90) are read once into a pinned int16 staging
    buffer and decoded on the device (wire.py); anything else goes through io.load_wav_file and is grouped by sample
    rate.  One peak search per channel serves decay, rt60bands and waterfall (pipeline.run_hot_path).  Files are
    rendered in order; the first file with a channel the reference would have raised on raises that ValueError after
    the reports of the files before it have been written (the reference's loop stops at the same point,
    bundle.py:56-67)."""
    from . import wire

    settings = settings or ReportSettings()
    decay_s, bands_s, water_s = resolved_settings(settings)
    spec_s = _apply_common_overrides(settings.spectrogram_analysis_settings or SpectrogramAnalysisSettings(), settings)
    modal_s = _apply_common_overrides(settings.modal_cloud_analysis_settings or ModalCloudAnalysisSettings(), settings)
    paths = [Path(p) for p in input_wav_file_paths]
    bases = [Path(b) for b in output_basenames]
    if not paths:
        return []
    mono = bool(settings.common_use_mono_downmix_for_stereo)
    groups: List[Tuple[List[_FileJob], engine.Rows, int, List[str]]] = []
    staged = wire.stage_pcm16(paths) if loaded is None else None
    if staged is not None and staged.channels in (1, 2):
        if staged.sample_rate_hz != settings.expected_sample_rate_hz:   # io.py:161-165
            raise ValueError(f"Expected sample rate {settings.expected_sample_rate_hz} Hz, "
                             f"but got {staged.sample_rate_hz} Hz for file {paths[0]}")
        import torch

        dev = engine._device(None)
        pcm_dev = staged.pcm.to(dev, non_blocking=True)
        rows = engine.pcm16_to_rows(pcm_dev, staged.frames, staged.channels, mono)
        per = _channel_names(staged.channels, mono)
        jobs = [_FileJob(p, b, (staged.frames, staged.channels), staged.sample_rate_hz, list(per), i * len(per))
                for i, (p, b) in enumerate(zip(paths, bases))]
        groups.append((jobs, rows, staged.sample_rate_hz, [n for _ in jobs for n in per]))
    else:
        by_rate: Dict[int, Tuple[List[_FileJob], List[np.ndarray], List[str]]] = {}
        for i, (p, b) in enumerate(zip(paths, bases)):
            audio = loaded[i] if loaded is not None else load_wav_file(
                p, expected_sample_rate_hz=settings.expected_sample_rate_hz,
                expected_channel_mode="mono_or_stereo", allow_mono_and_upmix_to_stereo=False)
            chans = get_analysis_channels(audio, mono)
            jobs, channels, names = by_rate.setdefault(int(audio.sample_rate_hz), ([], [], []))
            jobs.append(_FileJob(p, b, tuple(int(v) for v in audio.samples.shape), int(audio.sample_rate_hz),
                                 [n for n, _ in chans], first_row=len(channels)))
            for n, c in chans:
                channels.append(c)
                names.append(n)
        for sr, (jobs, channels, names) in by_rate.items():
            same_len = len({c.size for c in channels}) == 1
            groups.append((jobs, engine.make_rows(np.stack(channels, axis=0) if same_len else channels), sr, names))
    rendered: Dict[Path, ReportResults] = {}
    failed: Dict[Path, Exception] = {}
    for jobs, rows, sr, names in groups:
        res = _analyse_group(rows, sr, names, settings, decay_s, bands_s, water_s, spec_s, modal_s)
        for job in jobs:
            sl = slice(job.first_row, job.first_row + len(job.names))
            blocks = {k: (None if v is None else v[sl]) for k, v in res.items()}
            err = None
            for key, msg in (("decay", "Not enough samples after trimming/ignoring to compute EDC."),
                             ("water", "Not enough samples for STFT after trimming/ignoring (need at least n_fft).")):
                if blocks[key] is not None and any(r is None for r in blocks[key]):
                    err = err or ValueError(msg)
            for key in ("spec", "modal"):
                if blocks[key] is not None:
                    err = err or next((r for r in blocks[key] if isinstance(r, Exception)), None)
            if err is not None:
                failed[job.path] = err
                continue
            rendered[job.path] = (job, blocks)
    out: List[ReportResults] = []
    for p in paths:   # the reference's order: write until the first file that fails
        if p in failed:
            raise failed[p]
        job, blocks = rendered[p]
        out.append(_render(job, settings, decay_s, bands_s, water_s, blocks["decay"], blocks["bands"], blocks["water"],
                           spec_s, blocks["spec"], modal_s, blocks["modal"]))
    return out


def run_report_from_wav_file(
    input_wav_file_path: str | Path,
    output_basename: str | Path,
    settings: Optional[ReportSettings] = None,
) -> ReportResults:
    """Single-file form with the reference's signature (report.py:222-398)."""
    return run_reports_batch([input_wav_file_path], [output_basename], settings)[0]
