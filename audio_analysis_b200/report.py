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
    audio: LoadedAudio
    names: List[str]           # analysed channel names
    first_row: int = 0         # index of this file's first channel in the batch


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
    shape = (int(job.audio.samples.shape[0]), max(2, int(job.audio.samples.shape[1])))
    parts = [_header(job.path, shape, job.audio.sample_rate_hz, settings.expected_sample_rate_hz)]
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


def run_reports_batch(
    input_wav_file_paths: Sequence[str | Path],
    output_basenames: Sequence[str | Path],
    settings: Optional[ReportSettings] = None,
    loaded: Optional[Sequence[LoadedAudio]] = None,
) -> List[ReportResults]:
    """Reports for many WAV files with every channel analysed in shared launches."""
    settings = settings or ReportSettings()
    decay_s, bands_s, water_s = resolved_settings(settings)
    jobs: List[_FileJob] = []
    channels: List[np.ndarray] = []
    names: List[str] = []
    for i, (path, base) in enumerate(zip(input_wav_file_paths, output_basenames)):
        audio = loaded[i] if loaded is not None else load_wav_file(
            path, expected_sample_rate_hz=settings.expected_sample_rate_hz,
            expected_channel_mode="mono_or_stereo", allow_mono_and_upmix_to_stereo=False)
        chans = get_analysis_channels(audio, settings.common_use_mono_downmix_for_stereo)
        job = _FileJob(Path(path), Path(base), audio, [n for n, _ in chans], first_row=len(channels))
        jobs.append(job)
        for n, c in chans:
            channels.append(c)
            names.append(n)
    if not jobs:
        return []
    sr = jobs[0].audio.sample_rate_hz
    same_len = len({c.size for c in channels}) == 1
    rows = engine.make_rows(np.stack(channels, axis=0) if same_len else channels)
    decay_r = analyse_decay_batch(rows, sr, names, decay_s) if settings.run_decay else None
    bands_r = analyse_rt60_bands_batch(rows, sr, names, bands_s) if settings.run_rt60_bands else None
    water_r = analyse_waterfall_batch(rows, sr, names, water_s) if settings.run_waterfall else None
    spec_s = _apply_common_overrides(settings.spectrogram_analysis_settings or SpectrogramAnalysisSettings(), settings)
    modal_s = _apply_common_overrides(settings.modal_cloud_analysis_settings or ModalCloudAnalysisSettings(), settings)
    spec_r = analyse_spectrogram_batch(rows, sr, names, spec_s) if settings.run_spectrogram else None
    modal_r = analyse_modal_cloud_batch(rows, sr, names, modal_s) if settings.run_modal_cloud else None
    out = []
    for job in jobs:
        sl = slice(job.first_row, job.first_row + len(job.names))
        out.append(_render(job, settings, decay_s, bands_s, water_s,
                           None if decay_r is None else decay_r[sl],
                           None if bands_r is None else bands_r[sl],
                           None if water_r is None else water_r[sl],
                           spec_s, None if spec_r is None else spec_r[sl],
                           modal_s, None if modal_r is None else modal_r[sl]))
    return out


def run_report_from_wav_file(
    input_wav_file_path: str | Path,
    output_basename: str | Path,
    settings: Optional[ReportSettings] = None,
) -> ReportResults:
    """Single-file form with the reference's signature (report.py:222-398)."""
    return run_reports_batch([input_wav_file_path], [output_basename], settings)[0]
