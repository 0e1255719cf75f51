"""Command line for the accelerated path: `decay`, `rt60bands`, `waterfall`, `bundle`, `deconvolve`, `spectrogram`,
`modalcloud`.

    python -m analyse.cli decay --input ir.wav --no_show
    python -m analyse.cli rt60bands --input ir.wav --mono --band_mode third --no_show
    python -m analyse.cli waterfall --input ir.wav --no_show
    python -m analyse.cli bundle --input analysis_runs/20250101_120000
    python -m analyse.cli deconvolve --recorded_wav_file_path capture.wav --sweep_wav_file_path sweep.wav
    python -m analyse.cli spectrogram --input ir.wav --no_show
    python -m analyse.cli modalcloud --input ir.wav --no_show

Argument names, defaults and quirks follow the reference's `analyse/cli.py`
(decay :461-531, rt60bands :536-636, waterfall :1010-1100, bundle :377-388, deconvolve :392-455 / :1219-1246,
spectrogram :842-917, modalcloud :1105-1145), including
`decay --compute_edt` defaulting to true, `rt60bands --trim_to_peak` being impossible to
switch off, and the tri-state `rt60bands --legend_values`.  Each command prints the same
summary text as the reference.  The reference's other seven subcommands are outside this
build and are rejected by argparse.
"""

from __future__ import annotations

import argparse
from pathlib import Path
from typing import Optional, Sequence

from .bundle import BundleRunSettings, run_bundle_report
from .deconvolve import DeconvolveSettings, default_output_ir_path, deconvolve_from_wav_files
from .decay import DecayAnalysisSettings, DecayPlotSettings, plot_decay_from_wav_file, summarise_decay_results_text
from .modalcloud import (ModalCloudAnalysisSettings, ModalCloudPlotSettings, plot_modal_cloud_from_wav_file,
                         summarise_modal_cloud_results_text)
from .rt60bands import (Rt60BandsAnalysisSettings, Rt60BandsPlotSettings, plot_rt60_bands_from_wav_file,
                        summarise_rt60_bands_results_text)
from .spectrogram import (SpectrogramAnalysisSettings, SpectrogramPlotSettings, plot_spectrogram_from_wav_file,
                          summarise_spectrogram_results_text)
from .waterfall import (WaterfallAnalysisSettings, WaterfallPlotSettings, plot_waterfall_from_wav_file,
                        summarise_waterfall_results_text)

BoolOpt = argparse.BooleanOptionalAction


def _io_arguments(p: argparse.ArgumentParser, png_hint: str) -> None:
    p.add_argument("--input", dest="input_wav_file_path", type=str, required=True,
                   help="Path to input WAV file (mono or stereo, 48 kHz expected).")
    p.add_argument("--output", dest="output_basename", type=str, default=None, help=f"If provided, saves {png_hint}")
    p.add_argument("--no_show", action="store_true",
                   help="If set, do not display plots interactively (useful when saving files).")


def build_parser() -> argparse.ArgumentParser:
    top = argparse.ArgumentParser(prog="analyse",
                                  description="Offline analysis tools for reverb outputs (B200 decay path).")
    sub = top.add_subparsers(dest="command_name", required=True, help="Analysis to run. Use: analyse <command> --help")

    bundle = sub.add_parser("bundle", help="Analyse an IR bundle folder (meta.json + taps/*.wav), per-tap reports.")
    bundle.add_argument("--input", dest="bundle_root", type=str, required=True, help="Bundle root folder")
    bundle.add_argument("--reports-subdir", dest="reports_subdir", type=str, default="reports",
                        help="Subfolder under bundle root to write outputs (default: reports)")

    dc = sub.add_parser("deconvolve", help="Deconvolve recorded sweep output into an impulse response WAV.")
    dc.add_argument("--recorded_wav_file_path", type=str, required=True,
                    help="Recorded output WAV (mono or stereo, 48 kHz expected).")
    dc.add_argument("--sweep_wav_file_path", type=str, required=True,
                    help="Original sweep WAV that was played (mono or stereo, 48 kHz expected).")
    dc.add_argument("--output_ir_wav_file_path", type=str, default=None,
                    help="Output IR WAV path. If omitted, writes next to recorded file as: <recorded_stem>_ir.wav")
    dc.add_argument("--regularization_relative", type=float, default=1e-10,
                    help="Regularization as fraction of max(|X|^2). Default: 1e-10.")
    dc.add_argument("--normalise_peak", action=BoolOpt, default=True,
                    help="If true, peak-normalise the output IR. Default: true.")
    dc.add_argument("--target_peak", type=float, default=0.95,
                    help="Target peak value if normalise_peak is enabled. Default: 0.95.")
    dc.add_argument("--remove_dc", action=BoolOpt, default=True,
                    help="If true, remove DC offset from the output IR. Default: true.")
    dc.add_argument("--output_length_mode", type=str, choices=["recorded", "full_fft"], default="recorded",
                    help="Length of output IR. Default: recorded.")

    decay = sub.add_parser("decay", help="Schroeder EDC + T20/T30/RT60 decay estimation")
    _io_arguments(decay, "a PNG: <basename>_decay.png")
    decay.add_argument("--trim_to_peak", action=BoolOpt, default=True,
                       help="If true, start analysis at the absolute peak sample (default: true).")
    decay.add_argument("--ignore-leading", dest="ignore_leading_seconds", type=float, default=0.0)
    decay.add_argument("--edc_floor_db", type=float, default=-120.0)
    decay.add_argument("--fit_lower_limit_db", type=float, default=-80.0)
    decay.add_argument("--smoothing", dest="edc_smoothing_window_samples", type=int, default=0)
    decay.add_argument("--mono", dest="use_mono_downmix", action="store_true", default=False)
    decay.add_argument("--compute_edt", action=BoolOpt, default=True,
                       help="If set, also compute EDT (fit 0..-10 dB and extrapolate).")

    bands = sub.add_parser("rt60bands", help="Band-limited RT60: Low/Mid/High T30 by default (optional T20/EDT).")
    _io_arguments(bands, "one PNG: <basename>_rt60bands.png")
    bands.add_argument("--band_mode", type=str, default="three", choices=["three", "octave", "third"],
                       help="Band mode: three (Low/Mid/High), octave, or third (1/3-octave). Default: three.")
    bands.add_argument("--f_min_hz", type=float, default=31.5,
                       help="Minimum band centre frequency for octave/third modes (Hz).")
    bands.add_argument("--f_max_hz", type=float, default=16000.0,
                       help="Maximum band centre frequency for octave/third modes (Hz).")
    bands.add_argument("--legend_values", action=BoolOpt, default=None,
                       help="Per-band numeric values in the legend. Default: on for three, off otherwise.")
    bands.add_argument("--low_upper_hz", type=float, default=250.0, help="Low band upper cutoff (Hz).")
    bands.add_argument("--mid_center_hz", type=float, default=1000.0, help="Mid band centre (Hz).")
    bands.add_argument("--mid_width_octaves", type=float, default=2.0, help="Mid band width in octaves.")
    bands.add_argument("--high_lower_hz", type=float, default=4000.0, help="High band lower cutoff (Hz).")
    bands.add_argument("--transition_width_octaves", type=float, default=(1.0 / 6.0),
                       help="FFT mask transition width (octaves). Default: 1/6 octave.")
    bands.add_argument("--include_t20", action="store_true", help="Also compute T20-derived RT60 per band.")
    bands.add_argument("--include_edt", action="store_true", help="Also compute EDT per band.")
    bands.add_argument("--mono", dest="use_mono_downmix", action="store_true",
                       help="If set, downmix stereo to mono before analysis.")
    bands.add_argument("--trim_to_peak", action="store_true", default=True)  # cannot be disabled (reference :632)
    bands.add_argument("--ignore-leading", dest="ignore_leading_seconds", type=float, default=0.0)
    bands.add_argument("--edc_floor_db", type=float, default=-120.0)
    bands.add_argument("--fit_lower_limit_db", type=float, default=-80.0)
    bands.add_argument("--smoothing", dest="edc_smoothing_window_samples", type=int, default=0)
    bands.add_argument("--filterbank", type=str, choices=["fft_mask", "butterworth"], default="fft_mask",
                       help="Extension of the B200 build: 'fft_mask' = the reference's circular FFT mask (default), "
                            "'butterworth' = zero-phase Butterworth sections (not the reference's filter).")
    bands.add_argument("--butterworth_order", type=int, default=4, help="Order of the Butterworth prototype (default 4).")

    wf = sub.add_parser("waterfall", help="Waterfall (CSD-style): spectral slices over time (3D default, 2D ridges).")
    _io_arguments(wf, "PNG(s): <basename>_waterfall_<CH>.png")
    wf.add_argument("--mono", dest="use_mono_downmix", action="store_true",
                    help="If set, downmix stereo to mono before analysis.")
    wf.add_argument("--trim_to_peak", action=BoolOpt, default=True,
                    help="If true, start analysis at the absolute peak sample (default: true).")
    wf.add_argument("--ignore-leading", dest="ignore_leading_seconds", type=float, default=0.0)
    wf.add_argument("--duration", dest="analysis_duration_seconds", type=float, default=None)
    wf.add_argument("--n_fft", type=int, default=4096)
    wf.add_argument("--hop_length", type=int, default=512)
    wf.add_argument("--no_hann_window", action="store_true")
    wf.add_argument("--f_min_hz", type=float, default=20.0)
    wf.add_argument("--f_max_hz", type=float, default=20000.0)
    wf.add_argument("--style", type=str, choices=["3d", "2d"], default="3d",
                    help="Plot style: 3d waterfall (default) or 2d stacked ridges.")
    wf.add_argument("--slice_mode", type=str, choices=["auto", "uniform_time", "uniform_frames"], default="auto",
                    help="How to choose slice times/frames.")
    wf.add_argument("--num_slices", type=int, default=18)
    wf.add_argument("--slice_spacing_seconds", type=float, default=0.05)
    wf.add_argument("--start_time_seconds", type=float, default=0.0)
    wf.add_argument("--end_time_seconds", type=float, default=None)
    wf.add_argument("--db_reference", type=str, choices=["global_max", "slice_max"], default="global_max",
                    help="Normalisation: global_max keeps absolute decay; slice_max shows shape only.")
    wf.add_argument("--dynamic_range_db", type=float, default=80.0)
    wf.add_argument("--floor_db", type=float, default=-120.0)
    wf.add_argument("--smoothing_log_bins", type=int, default=0)
    wf.add_argument("--log_bins_per_octave", type=int, default=96)
    wf.add_argument("--elev_deg", type=float, default=30.0)
    wf.add_argument("--azim_deg", type=float, default=-60.0)
    wf.add_argument("--ridge_offset_db", type=float, default=6.0)

    sp = sub.add_parser("spectrogram", help="Plot time–frequency magnitude spectrogram (log-frequency).")
    _io_arguments(sp, "PNG(s): <basename>_spectrogram_<CH>.png")
    sp.add_argument("--mono", dest="use_mono_downmix", action="store_true",
                    help="If set, downmix stereo to mono before analysis.")
    sp.add_argument("--trim_to_peak", action=BoolOpt, default=True,
                    help="If true, start analysis at the absolute peak sample (default: true).")
    sp.add_argument("--ignore-leading", dest="ignore_leading_seconds", type=float, default=0.0,
                    help="Ignore this many seconds after trim point (default: 0.0).")
    sp.add_argument("--duration", dest="analysis_duration_seconds", type=float, default=None,
                    help="If set, analyse only this many seconds from the analysed segment.")
    sp.add_argument("--n_fft", type=int, default=4096, help="STFT FFT size (default: 4096).")
    sp.add_argument("--hop_length", type=int, default=512, help="STFT hop length (default: 512).")
    sp.add_argument("--no_hann_window", action="store_true", help="If set, do not apply a Hann window to each frame.")
    sp.add_argument("--floor_db", type=float, default=-120.0, help="Magnitude floor in dB (default: -120).")
    sp.add_argument("--f_min_hz", type=float, default=20.0, help="Min frequency to display (default: 20).")
    sp.add_argument("--f_max_hz", type=float, default=20000.0, help="Max frequency to display (default: 20000).")
    sp.add_argument("--dynamic_range_db", type=float, default=90.0,
                    help="Color scale range below max (default: 90). Use 0 to disable and use percentiles.")

    mc = sub.add_parser("modalcloud", help="Modal cloud: frequency vs RT60 points from per-bin STFT decay fits.")
    mc.add_argument("--input", dest="input_wav_file_path", type=str, required=True)
    mc.add_argument("--output", dest="output_basename", type=str, default=None)
    mc.add_argument("--no_show", action="store_true")
    mc.add_argument("--mono", dest="use_mono_downmix", action="store_true")
    mc.add_argument("--trim_to_peak", action=BoolOpt, default=True)
    mc.add_argument("--ignore-leading", dest="ignore_leading_seconds", type=float, default=0.0)
    mc.add_argument("--duration", dest="analysis_duration_seconds", type=float, default=None)
    mc.add_argument("--n_fft", type=int, default=8192)
    mc.add_argument("--hop_length", type=int, default=512)
    mc.add_argument("--no_hann_window", action="store_true")
    mc.add_argument("--f_min_hz", type=float, default=20.0)
    mc.add_argument("--f_max_hz", type=float, default=20000.0)
    mc.add_argument("--metric", type=str, choices=["t30", "t20", "edt"], default="t30",
                    help="Which decay fit window to use per bin (default: t30).")
    mc.add_argument("--log_bins_per_octave", type=int, default=24)
    mc.add_argument("--min_bins", type=int, default=24)
    mc.add_argument("--fit_lower_limit_db", type=float, default=-80.0)
    mc.add_argument("--min_fit_points", type=int, default=10)
    mc.add_argument("--min_peak_db_above_floor", type=float, default=20.0)
    mc.add_argument("--floor_db", type=float, default=-120.0)
    mc.add_argument("--show_median_curve", action=BoolOpt, default=True)
    mc.add_argument("--median_octave_window", type=float, default=0.25)
    mc.add_argument("--ylim_seconds_min", type=float, default=None)
    mc.add_argument("--ylim_seconds_max", type=float, default=None)
    return top


def parse_arguments(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    return build_parser().parse_args(argv)


def _basename(args: argparse.Namespace) -> Optional[str]:
    return None if args.output_basename is None else str(Path(args.output_basename))


def decay_settings_from_args(a: argparse.Namespace, compute_edt: bool) -> DecayAnalysisSettings:
    return DecayAnalysisSettings(
        trim_to_peak=bool(a.trim_to_peak),
        ignore_leading_seconds=float(a.ignore_leading_seconds),
        edc_floor_db=float(a.edc_floor_db),
        fit_lower_limit_db=float(a.fit_lower_limit_db),
        edc_smoothing_window_samples=int(a.edc_smoothing_window_samples),
        use_mono_downmix_for_stereo=bool(a.use_mono_downmix),
        compute_edt=bool(compute_edt),
    )


def run_decay_command(a: argparse.Namespace) -> str:
    results = plot_decay_from_wav_file(
        input_wav_file_path=str(a.input_wav_file_path),
        analysis_settings=decay_settings_from_args(a, a.compute_edt),
        plot_settings=DecayPlotSettings(),
        output_basename=_basename(a),
        show_interactive=not bool(a.no_show),
    )
    return summarise_decay_results_text(results)


def run_rt60bands_command(a: argparse.Namespace) -> str:
    settings = Rt60BandsAnalysisSettings(
        band_mode=str(a.band_mode),
        low_upper_hz=float(a.low_upper_hz),
        mid_center_hz=float(a.mid_center_hz),
        mid_width_octaves=float(a.mid_width_octaves),
        high_lower_hz=float(a.high_lower_hz),
        f_min_hz=float(a.f_min_hz),
        f_max_hz=float(a.f_max_hz),
        transition_width_octaves=float(a.transition_width_octaves),
        include_t20=bool(a.include_t20),
        include_edt=bool(a.include_edt),
        decay_settings=decay_settings_from_args(a, a.include_edt),
        filterbank=str(a.filterbank),
        butterworth_order=int(a.butterworth_order),
    )
    legend = (str(a.band_mode) == "three") if a.legend_values is None else bool(a.legend_values)
    results = plot_rt60_bands_from_wav_file(
        input_wav_file_path=str(a.input_wav_file_path),
        settings=settings,
        plot_settings=Rt60BandsPlotSettings(legend_values=legend),
        output_basename=_basename(a),
        show_interactive=not bool(a.no_show),
    )
    return summarise_rt60_bands_results_text(results, include_t20=settings.include_t20, include_edt=settings.include_edt)


def run_waterfall_command(a: argparse.Namespace) -> str:
    analysis = WaterfallAnalysisSettings(
        use_mono_downmix_for_stereo=bool(a.use_mono_downmix),
        trim_to_peak=bool(a.trim_to_peak),
        ignore_leading_seconds=float(a.ignore_leading_seconds),
        analysis_duration_seconds=a.analysis_duration_seconds,
        n_fft=int(a.n_fft),
        hop_length=int(a.hop_length),
        use_hann_window=not bool(a.no_hann_window),
        f_min_hz=float(a.f_min_hz),
        f_max_hz=float(a.f_max_hz),
        slice_mode=str(a.slice_mode),
        num_slices=int(a.num_slices),
        slice_spacing_seconds=float(a.slice_spacing_seconds),
        start_time_seconds=float(a.start_time_seconds),
        end_time_seconds=a.end_time_seconds,
        db_reference=str(a.db_reference),
        smoothing_log_bins=int(a.smoothing_log_bins),
        log_bins_per_octave=int(a.log_bins_per_octave),
        dynamic_range_db=float(a.dynamic_range_db),
        floor_db=float(a.floor_db),
    )
    plot = WaterfallPlotSettings(style=str(a.style), elev_deg=float(a.elev_deg), azim_deg=float(a.azim_deg),
                                 ridge_offset_db=float(a.ridge_offset_db))
    results = plot_waterfall_from_wav_file(
        input_wav_file_path=str(a.input_wav_file_path),
        analysis_settings=analysis,
        plot_settings=plot,
        output_basename=_basename(a),
        show_interactive=not bool(a.no_show),
    )
    return summarise_waterfall_results_text(results)


def run_spectrogram_command(a: argparse.Namespace) -> str:
    dyn = float(a.dynamic_range_db)
    analysis = SpectrogramAnalysisSettings(
        use_mono_downmix_for_stereo=bool(a.use_mono_downmix),
        trim_to_peak=bool(a.trim_to_peak),
        ignore_leading_seconds=float(a.ignore_leading_seconds),
        analysis_duration_seconds=a.analysis_duration_seconds,
        n_fft=int(a.n_fft),
        hop_length=int(a.hop_length),
        use_hann_window=not bool(a.no_hann_window),
        floor_db=float(a.floor_db),
        f_min_hz=float(a.f_min_hz),
        f_max_hz=float(a.f_max_hz),
        dynamic_range_db=None if dyn <= 0.0 else dyn,   # reference cli.py:1402-1406
    )
    results = plot_spectrogram_from_wav_file(
        input_wav_file_path=str(a.input_wav_file_path),
        analysis_settings=analysis,
        plot_settings=SpectrogramPlotSettings(),
        output_basename=_basename(a),
        show_interactive=not bool(a.no_show),
    )
    return summarise_spectrogram_results_text(results)


def run_modalcloud_command(a: argparse.Namespace) -> str:
    analysis = ModalCloudAnalysisSettings(
        use_mono_downmix_for_stereo=bool(a.use_mono_downmix),
        trim_to_peak=bool(a.trim_to_peak),
        ignore_leading_seconds=float(a.ignore_leading_seconds),
        analysis_duration_seconds=a.analysis_duration_seconds,
        n_fft=int(a.n_fft),
        hop_length=int(a.hop_length),
        use_hann_window=not bool(a.no_hann_window),
        f_min_hz=float(a.f_min_hz),
        f_max_hz=float(a.f_max_hz),
        log_bins_per_octave=int(a.log_bins_per_octave),
        min_bins=int(a.min_bins),
        metric=str(a.metric),
        fit_lower_limit_db=float(a.fit_lower_limit_db),
        min_fit_points=int(a.min_fit_points),
        min_peak_db_above_floor=float(a.min_peak_db_above_floor),
        floor_db=float(a.floor_db),
    )
    ylim = None
    if a.ylim_seconds_min is not None and a.ylim_seconds_max is not None:
        ylim = (float(a.ylim_seconds_min), float(a.ylim_seconds_max))
    plot = ModalCloudPlotSettings(show_median_curve=bool(a.show_median_curve),
                                  median_octave_window=float(a.median_octave_window), ylim_seconds=ylim)
    results = plot_modal_cloud_from_wav_file(
        input_wav_file_path=str(a.input_wav_file_path),
        analysis_settings=analysis,
        plot_settings=plot,
        output_basename=_basename(a),
        show_interactive=not bool(a.no_show),
    )
    return summarise_modal_cloud_results_text(results)


def run_deconvolve_command(a: argparse.Namespace) -> str:
    """Reference cli.py:1219-1246: same settings mapping and the same confirmation lines."""
    out = str(Path(a.output_ir_wav_file_path)) if a.output_ir_wav_file_path is not None \
        else str(default_output_ir_path(a.recorded_wav_file_path))
    settings = DeconvolveSettings(
        regularization_relative=float(a.regularization_relative),
        normalise_peak=bool(a.normalise_peak),
        target_peak=float(a.target_peak),
        remove_dc=bool(a.remove_dc),
        output_length_mode=str(a.output_length_mode),
    )
    result = deconvolve_from_wav_files(str(a.recorded_wav_file_path), str(a.sweep_wav_file_path), settings, out)
    return "\n".join([
        f"Wrote IR WAV: {out}",
        f"  sample_rate_hz={result.sample_rate_hz}",
        f"  channels={result.samples.shape[1]}",
        f"  length_seconds={result.samples.shape[0] / float(result.sample_rate_hz):.3f}",
    ])


def run_bundle_command(a: argparse.Namespace) -> str:
    index = run_bundle_report(str(a.bundle_root), settings=BundleRunSettings(reports_subdir=str(a.reports_subdir)))
    return f"Wrote bundle report index: {index}"


COMMANDS = {
    "decay": run_decay_command,
    "rt60bands": run_rt60bands_command,
    "waterfall": run_waterfall_command,
    "bundle": run_bundle_command,
    "deconvolve": run_deconvolve_command,
    "spectrogram": run_spectrogram_command,
    "modalcloud": run_modalcloud_command,
}


def main(argv: Optional[Sequence[str]] = None) -> None:
    args = parse_arguments(argv)
    handler = COMMANDS.get(str(args.command_name))
    if handler is None:
        raise ValueError(f"Unknown command: {args.command_name}")
    print(handler(args))


if __name__ == "__main__":
    main()
