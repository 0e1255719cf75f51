"""Waterfall / cumulative spectral decay slices on the B200 path.

Drop-in mirror of the reference's `analyse/waterfall.py` API.  Host side: which
STFT frames become slices (a function of the frame count and the settings only,
reference waterfall.py:233-286) and which rfft bins are kept (:300-308).  Device
side (`aa_csd_batch`): windowed real FFT of the selected frames, magnitude in dB,
crop, reference level and clip (reference :188-230, :309, :328-339).  The
reference computes every frame and keeps <= num_slices of them; the result is
the same because normalisation only looks at the selected slices (:335).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine
from . import _native as nat
from .io import get_analysis_channels, load_wav_file


@dataclass(frozen=True)
class WaterfallAnalysisSettings:
    use_mono_downmix_for_stereo: bool = False
    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    analysis_duration_seconds: Optional[float] = None
    n_fft: int = 4096
    hop_length: int = 512
    use_hann_window: bool = True
    f_min_hz: float = 20.0
    f_max_hz: float = 20000.0
    slice_mode: str = "auto"  # "auto" | "uniform_time" | "uniform_frames"
    num_slices: int = 18
    slice_spacing_seconds: float = 0.05
    start_time_seconds: float = 0.0
    end_time_seconds: Optional[float] = None
    db_reference: str = "global_max"  # "global_max" | "slice_max"
    smoothing_log_bins: int = 0
    log_bins_per_octave: int = 96
    dynamic_range_db: float = 80.0
    floor_db: float = -120.0


@dataclass(frozen=True)
class WaterfallPlotSettings:
    style: str = "3d"  # "3d" | "2d"
    secondary_channel_alpha: float = 0.7
    elev_deg: float = 30.0
    azim_deg: float = -60.0
    ridge_offset_db: float = 6.0
    zlim_db: Optional[Tuple[float, float]] = None


@dataclass(frozen=True)
class ChannelWaterfallResult:
    channel_name: str
    sample_rate_hz: int
    analysis_start_sample_index: int
    analysis_length_samples: int
    slice_times_seconds: np.ndarray      # (S,)
    frequency_hz: np.ndarray             # (F,)
    slice_magnitude_rel_db: np.ndarray   # (S, F) in [-dynamic_range_db, 0]


_TOO_SHORT = "Not enough samples after trimming/selection for waterfall (need at least n_fft)."
_TOO_FEW_SLICES = "Not enough slices selected for waterfall (increase duration or num_slices)."


# --------------------------------------------------------------------------------------
# host-side selection logic
# --------------------------------------------------------------------------------------


def frame_times_seconds(num_frames: int, hop_length: int, sample_rate_hz: int) -> np.ndarray:
    """Window-start time of every STFT frame, float32 like the reference (waterfall.py:229)."""
    return (np.arange(int(num_frames), dtype=np.float32) * float(hop_length) / float(sample_rate_hz)).astype(np.float32)


def _select_slice_frame_indices(frame_times: np.ndarray, settings: WaterfallAnalysisSettings) -> np.ndarray:
    """Frame index of every slice, ordered and unique (reference waterfall.py:233-286)."""
    if frame_times.size == 0:
        return np.zeros((0,), dtype=np.int32)
    t_begin = float(max(0.0, settings.start_time_seconds))
    last = float(frame_times[-1])
    t_end = last if settings.end_time_seconds is None else float(settings.end_time_seconds)
    if t_end <= t_begin:
        t_end = last
    usable = np.nonzero((frame_times >= t_begin) & (frame_times <= t_end))[0]
    if usable.size == 0:
        return np.zeros((0,), dtype=np.int32)
    first, final = int(usable[0]), int(usable[-1])
    mode = str(settings.slice_mode).lower()
    if mode == "uniform_frames":
        return np.unique(np.linspace(first, final, int(max(1, settings.num_slices)), dtype=np.int32))
    if mode == "uniform_time":
        spacing = float(max(1e-4, settings.slice_spacing_seconds))
        wanted = np.arange(t_begin, t_end + 1e-9, spacing, dtype=np.float64)
    else:
        wanted = np.linspace(t_begin, t_end, int(max(2, settings.num_slices)), dtype=np.float64)
    nearest = [int(np.argmin(np.abs(frame_times - float(t)))) for t in wanted]
    picked = [j for j in nearest if first <= j <= final]
    if mode == "uniform_time" and not picked:
        picked = [first, final]
    return np.unique(np.array(picked, dtype=np.int32))


def frequency_selection(settings: WaterfallAnalysisSettings, sample_rate_hz: int) -> Tuple[int, np.ndarray]:
    """(first kept rfft bin, kept frequencies float32) - reference waterfall.py:213, :300-308."""
    freq = np.fft.rfftfreq(int(settings.n_fft), d=1.0 / float(sample_rate_hz)).astype(np.float32)
    nyquist = float(freq[-1]) if freq.size else 0.0
    lo = float(np.clip(settings.f_min_hz, 1.0, nyquist))
    hi = float(np.clip(settings.f_max_hz, lo, nyquist))
    keep = np.nonzero((freq >= lo) & (freq <= hi))[0]
    if keep.size == 0:
        raise ValueError("Waterfall frequency selection is empty (check f_min_hz/f_max_hz).")
    return int(keep[0]), freq[keep[0]: keep[-1] + 1].astype(np.float32)


def analysed_span(length: int, peak: int, settings: WaterfallAnalysisSettings, sample_rate_hz: int) -> Tuple[int, int]:
    """(start index, analysed length) after trim / ignore / duration (reference waterfall.py:355-372)."""
    start = int(peak) if settings.trim_to_peak else 0
    remaining = int(length) - start
    if settings.ignore_leading_seconds > 0.0:
        skip = int(round(float(settings.ignore_leading_seconds) * float(sample_rate_hz)))
        skip = max(0, min(skip, remaining))
        start += skip
        remaining -= skip
    if settings.analysis_duration_seconds is not None:
        keep = int(round(float(settings.analysis_duration_seconds) * float(sample_rate_hz)))
        remaining = max(0, min(keep, remaining))
    return start, remaining


# --------------------------------------------------------------------------------------
# optional log-frequency smoothing of finished slices (reference waterfall.py:132-185); off by
# default.  It post-processes the S x F dB table (<= 18 x 1705 values per channel), so it runs on
# the host on the kernel's raw dB output and the normalisation is redone afterwards.
# --------------------------------------------------------------------------------------


def smoothing_tables(frequency_hz: np.ndarray, settings: WaterfallAnalysisSettings):
    """Host side of the log-frequency smoothing (reference waterfall.py:159-176): which kept bins lie inside
    [f_min, f_max], the uniform log2(f) grid, and for every grid point / bin the interval of the other axis that
    np.interp would pick.  Returns None when no bin is inside (the slices then stay unsmoothed, :164-165), else
    (k0, f_in float64, grid float64, grid_seg int32, bin_seg int32).  The arithmetic on the curves is the kernel's."""
    f = frequency_hz.astype(np.float64, copy=False)
    lo = float(max(1.0, settings.f_min_hz))
    hi = float(max(lo, settings.f_max_hz))
    inside = np.flatnonzero((f >= lo) & (f <= hi))
    if inside.size == 0:
        return None
    k0 = int(inside[0])
    f_in = f[k0:k0 + inside.size]
    a, b = float(np.log2(f_in[0])), float(np.log2(f_in[-1]))
    per_octave = int(max(16, settings.log_bins_per_octave))
    count = int(max(8, np.ceil((b - a) * per_octave))) + 1
    grid = 2.0 ** np.linspace(a, b, count, dtype=np.float64)

    def segments(x, xp):
        seg = (np.searchsorted(xp, x, side="right") - 1).astype(np.int32)
        seg[x < xp[0]] = -1
        seg[x >= xp[-1]] = -2
        return seg

    return k0, f_in, grid, segments(grid, f_in), segments(f_in, grid)


# --------------------------------------------------------------------------------------
# batched analysis
# --------------------------------------------------------------------------------------


class WaterfallPlan:
    """Per-row selection for a batch: analysis start / length and the slice frames.

    The selection is computed on the device; the host copies (`start`, `length`, `frames`, `n_slices`
    as numpy arrays) are made on first access, so enqueueing a batch never waits for the GPU.
    """

    def __init__(self, start, length, frames, n_slices, bin_lo: int, freq: np.ndarray, hop_length: int,
                 sample_rate_hz: int, n_fft: int = 0):
        self.n_fft = int(n_fft)
        self._dev = {"start": start, "length": length, "frames": frames, "n_slices": n_slices}
        self._host: Dict[str, np.ndarray] = {}
        self.bin_lo = int(bin_lo)
        self.freq = freq                      # [F] float32
        self.hop_length = int(hop_length)
        self.sample_rate_hz = int(sample_rate_hz)

    def _get(self, name: str, dtype) -> np.ndarray:
        if name not in self._host:
            v = self._dev[name]
            self._host[name] = (v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)).astype(dtype)
        return self._host[name]

    @property
    def start(self) -> np.ndarray:        # [B] int32
        return self._get("start", np.int32)

    @property
    def length(self) -> np.ndarray:       # [B] int64 analysed samples
        return self._get("length", np.int64)

    @property
    def frames(self) -> np.ndarray:       # [B, S_max] int32
        return self._get("frames", np.int32)

    @property
    def n_slices(self) -> np.ndarray:     # [B] int32
        return self._get("n_slices", np.int32)

    def slice_times(self, row: int) -> np.ndarray:
        """Window-start times of the row's slices, float32 like the reference (waterfall.py:229, :400)."""
        idx = self.frames[row, : int(self.n_slices[row])]
        return (idx.astype(np.float32) * float(self.hop_length) / float(self.sample_rate_hz)).astype(np.float32)


_slice_tables: Dict[tuple, Tuple[np.ndarray, np.ndarray]] = {}


def slice_table(settings: WaterfallAnalysisSettings, sample_rate_hz: int, max_frames: int):
    """frames[t] / counts[t] for every possible frame count t in [0, max_frames].

    The selected slices depend only on the number of STFT frames and the settings
    (reference waterfall.py:233-286), so the whole mapping is tabulated once per
    (settings, sample rate) and looked up per row - on the device for batches.
    """
    key = (settings.slice_mode, int(settings.num_slices), float(settings.slice_spacing_seconds),
           float(settings.start_time_seconds), settings.end_time_seconds, int(settings.hop_length),
           int(sample_rate_hz))
    have = _slice_tables.get(key)
    if have is not None and have[0].shape[0] > max_frames:
        return have
    rows: List[np.ndarray] = [np.zeros((0,), dtype=np.int32)]
    for t in range(1, int(max_frames) + 1):
        rows.append(_select_slice_frame_indices(frame_times_seconds(t, settings.hop_length, sample_rate_hz), settings))
    width = max(1, max(int(r.size) for r in rows))
    if width > nat.AA_MAX_SLICES:
        raise ValueError(f"more than {nat.AA_MAX_SLICES} slices selected")
    frames = np.zeros((len(rows), width), dtype=np.int32)
    counts = np.zeros(len(rows), dtype=np.int32)
    for t, r in enumerate(rows):
        frames[t, : r.size] = r
        counts[t] = r.size
    _slice_tables[key] = (frames, counts)
    return frames, counts


_device_tables: Dict[tuple, tuple] = {}


def _device_slice_table(table: np.ndarray, counts: np.ndarray, dev):
    """Device copies of a slice table, cached per (table identity, device)."""
    import torch

    key = (id(table), str(dev))
    have = _device_tables.get(key)
    if have is None or have[0] is not table:
        have = (table, torch.from_numpy(table).to(dev), torch.from_numpy(counts).to(dev))
        _device_tables[key] = have
    return have[1], have[2]


def run_waterfall(rows: engine.Rows, sample_rate_hz: int, settings: WaterfallAnalysisSettings, peaks=None,
                  validate: bool = True):
    """Device-level call: returns (slices tensor [B, S_max, F], WaterfallPlan).

    `peaks` (device int32 [B]) may be supplied when the peak index is already known from the decay
    kernel, which saves the 4 B/sample arg-max pass.  Start / span / slice lookup are
    integer bookkeeping on the device, so the CSD kernel is enqueued without a host round trip; the
    host copy of the plan (and the reference's ValueErrors) follows after the launch.
    """
    import torch

    n_fft, hop = int(settings.n_fft), int(settings.hop_length)
    if n_fft <= 0 or hop <= 0:
        raise ValueError("n_fft and hop_length must be positive.")
    bin_lo, freq = frequency_selection(settings, sample_rate_hz)
    dev = rows.x.device
    b = rows.n_rows
    peaks_d = None
    if settings.trim_to_peak:
        peaks_d = (peaks if peaks is not None else engine.peak_index(rows))
        if peaks_d.dtype != torch.int32:
            peaks_d = peaks_d.to(torch.int32)
    skip = 0
    if settings.ignore_leading_seconds > 0.0:  # reference waterfall.py:363-367
        skip = max(0, int(round(float(settings.ignore_leading_seconds) * float(sample_rate_hz))))
    keep = -1
    if settings.analysis_duration_seconds is not None:  # reference waterfall.py:369-372
        keep = max(0, int(round(float(settings.analysis_duration_seconds) * float(sample_rate_hz))))
    max_frames = max(0, 1 + (rows.n - n_fft) // hop)
    table, counts = slice_table(settings, sample_rate_hz, max_frames)
    table_d, counts_d = _device_slice_table(table, counts, dev)
    # start / span / frame count / the row's line of the slice table: one small kernel of the library
    start, remaining, frames_d, nslices_d = engine.csd_plan(rows, peaks_d, skip, keep, n_fft, hop, table_d, counts_d)
    tables = smoothing_tables(freq, settings) if int(settings.smoothing_log_bins or 0) > 1 else None
    params = engine.csd_params(n_fft, hop, settings.use_hann_window, bin_lo, int(freq.size), settings.db_reference,
                               settings.floor_db, settings.dynamic_range_db, raw_db=tables is not None)
    slices = engine.csd(rows, start, frames_d, nslices_d, params)
    if tables is not None:   # reference waterfall.py:312-339 on the device: smooth, then reference level and clip
        engine.csd_smooth(slices, nslices_d, *tables, int(settings.smoothing_log_bins), settings.db_reference,
                          settings.dynamic_range_db)
    plan = WaterfallPlan(start=start, length=remaining, frames=frames_d, n_slices=nslices_d, bin_lo=bin_lo, freq=freq,
                         hop_length=hop, sample_rate_hz=int(sample_rate_hz), n_fft=n_fft)
    if validate:
        if np.any(plan.length < n_fft):
            raise ValueError(_TOO_SHORT)
        if np.any(plan.n_slices < 2):
            raise ValueError(_TOO_FEW_SLICES)
    return slices, plan


def analyse_waterfall_batch(channels, sample_rate_hz: int, channel_names: Optional[Sequence[str]] = None,
                            settings: Optional[WaterfallAnalysisSettings] = None, device=None,
                            peaks=None) -> List[ChannelWaterfallResult]:
    settings = settings or WaterfallAnalysisSettings()
    rows = channels if isinstance(channels, engine.Rows) else engine.make_rows(channels, device)
    names = list(channel_names) if channel_names is not None else [f"ch{i}" for i in range(rows.n_rows)]
    slices, plan = run_waterfall(rows, sample_rate_hz, settings, peaks)
    return waterfall_results(slices, plan, names, sample_rate_hz)


def waterfall_results(slices, plan: WaterfallPlan, names: Sequence[str], sample_rate_hz: int,
                      strict: bool = True) -> List[Optional[ChannelWaterfallResult]]:
    """Result records from the device table.  strict=False: a row the reference would raise on (too short, fewer than
    two slices) yields None instead of an exception (batched reports isolate such files)."""
    table = slices.cpu().numpy()
    results: List[Optional[ChannelWaterfallResult]] = []
    for b in range(table.shape[0]):
        s = int(plan.n_slices[b])
        if s < 2 or int(plan.length[b]) < plan.n_fft:
            if strict:
                raise ValueError(_TOO_SHORT if int(plan.length[b]) < plan.n_fft else _TOO_FEW_SLICES)
            results.append(None)
            continue
        rel = np.array(table[b, :s, :], dtype=np.float32)
        results.append(ChannelWaterfallResult(
            channel_name=str(names[b]),
            sample_rate_hz=int(sample_rate_hz),
            analysis_start_sample_index=int(plan.start[b]),
            analysis_length_samples=int(plan.length[b]),
            slice_times_seconds=plan.slice_times(b),
            frequency_hz=plan.freq,
            slice_magnitude_rel_db=rel,
        ))
    return results


def analyse_waterfall_for_channel(
    samples: np.ndarray,
    sample_rate_hz: int,
    channel_name: str,
    settings: WaterfallAnalysisSettings,
) -> ChannelWaterfallResult:
    """One channel (reference waterfall.py:349-410)."""
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("_compute_stft_magnitude_db expects a 1D mono array.")
    if samples.size < int(settings.n_fft):
        raise ValueError(_TOO_SHORT)
    return analyse_waterfall_batch(samples[None, :], sample_rate_hz, [channel_name], settings)[0]


def analyse_waterfall_from_wav_file(
    input_wav_file_path: str | Path,
    settings: Optional[WaterfallAnalysisSettings] = None,
) -> List[ChannelWaterfallResult]:
    """Reference waterfall.py:413-441; the channels of one file share a batch."""
    settings = settings or WaterfallAnalysisSettings()
    audio = load_wav_file(input_wav_file_path, expected_channel_mode="mono_or_stereo",
                          allow_mono_and_upmix_to_stereo=False)
    chans = get_analysis_channels(audio, use_mono_downmix_for_stereo=settings.use_mono_downmix_for_stereo)
    data = np.stack([c for _, c in chans], axis=0)
    return analyse_waterfall_batch(data, audio.sample_rate_hz, [n for n, _ in chans], settings)


# --------------------------------------------------------------------------------------
# figures (optional) and text
# --------------------------------------------------------------------------------------

_TICKS_HZ = (20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000)


def _hz_major_ticks_for_audio(f_min_hz: float, f_max_hz: float) -> List[float]:
    inside = [float(t) for t in _TICKS_HZ if float(f_min_hz) <= float(t) <= float(f_max_hz)]
    return inside or [float(max(1.0, f_min_hz)), float(f_max_hz)]


def _hz_tick_formatter(x, pos) -> str:
    return f"{int(round(x / 1000.0))}k" if x >= 1000.0 else f"{int(round(x))}"


def plot_waterfall_figure(
    result: ChannelWaterfallResult,
    analysis_settings: WaterfallAnalysisSettings,
    plot_settings: WaterfallPlotSettings,
    title: Optional[str] = None,
):
    """3-D surface over log10(f) (default) or 2-D stacked ridges (same content as reference :458-560)."""
    from .plotting import DEFAULT_DPI, DEFAULT_FIGURE_SIZE, create_figure_and_axis, pyplot

    plt = pyplot(optional=False)
    import matplotlib.ticker as mticker

    nyquist = 0.5 * float(result.sample_rate_hz)
    f_lo = float(np.clip(analysis_settings.f_min_hz, 1.0, nyquist))
    f_hi = float(np.clip(analysis_settings.f_max_hz, f_lo, nyquist))
    dyn = float(max(10.0, analysis_settings.dynamic_range_db))
    ticks = _hz_major_ticks_for_audio(f_lo, f_hi)
    if str(plot_settings.style).lower() == "2d":
        figure, axis = create_figure_and_axis(title=title)
        axis.set_xlabel("Frequency (Hz)")
        axis.set_ylabel("Magnitude (dB, offset by time slice)")
        axis.set_xscale("log")
        axis.set_xlim(f_lo, f_hi)
        axis.set_xticks(ticks)
        axis.xaxis.set_major_formatter(mticker.FuncFormatter(_hz_tick_formatter))
        axis.xaxis.set_minor_formatter(mticker.NullFormatter())
        step = float(max(0.0, plot_settings.ridge_offset_db))
        count = int(result.slice_times_seconds.size)
        for i in range(count):
            axis.plot(result.frequency_hz, result.slice_magnitude_rel_db[i] - float(i) * step, alpha=0.9, label=None)
        for i in (0, count // 2, count - 1):
            axis.text(float(result.frequency_hz[0]), -float(i) * step, f"{float(result.slice_times_seconds[i]):.2f}s",
                      fontsize=9, verticalalignment="bottom")
        axis.grid(True, which="both", linestyle=":", linewidth=0.5)
        if plot_settings.zlim_db is not None:
            axis.set_ylim(*plot_settings.zlim_db)
        else:
            axis.set_ylim(-float(count - 1) * step - dyn, 2.0)
        return figure
    from mpl_toolkits.mplot3d import Axes3D  # noqa: F401  (registers the 3d projection)

    figure = plt.figure(figsize=DEFAULT_FIGURE_SIZE, dpi=DEFAULT_DPI)
    axis = figure.add_subplot(111, projection="3d")
    if title:
        axis.set_title(title)
    gx, gy = np.meshgrid(np.log10(result.frequency_hz.astype(np.float64)), result.slice_times_seconds.astype(np.float64))
    axis.plot_surface(gx, gy, result.slice_magnitude_rel_db.astype(np.float64), cmap="viridis", alpha=0.8,
                      antialiased=True, edgecolor="none", linewidth=0)
    axis.set_xlabel("Frequency (Hz)")
    axis.set_ylabel("Time (s)")
    axis.set_zlabel("Magnitude (dB rel)")
    axis.invert_yaxis()
    axis.set_xlim(np.log10(f_lo), np.log10(f_hi))
    axis.set_xticks([np.log10(t) for t in ticks])
    axis.set_xticklabels([_hz_tick_formatter(t, None) for t in ticks])
    if plot_settings.zlim_db is not None:
        axis.set_zlim(*plot_settings.zlim_db)
    else:
        axis.set_zlim(-dyn, 2.0)
    axis.view_init(elev=float(plot_settings.elev_deg), azim=float(plot_settings.azim_deg))
    return figure


def plot_waterfall_from_wav_file(
    input_wav_file_path: str | Path,
    analysis_settings: Optional[WaterfallAnalysisSettings] = None,
    plot_settings: Optional[WaterfallPlotSettings] = None,
    output_basename: Optional[str | Path] = None,
    show_interactive: bool = True,
) -> List[ChannelWaterfallResult]:
    """Analyse, then one `<basename>_waterfall_<CH>.png` per channel when matplotlib is available (:563-607)."""
    from .plotting import finalize_and_show_or_save, pyplot

    analysis_settings = analysis_settings or WaterfallAnalysisSettings()
    plot_settings = plot_settings or WaterfallPlotSettings()
    results = analyse_waterfall_from_wav_file(input_wav_file_path, analysis_settings)
    if (output_basename is not None or show_interactive) and pyplot() is not None:
        for r in results:
            figure = plot_waterfall_figure(r, analysis_settings, plot_settings,
                                           title=f"Waterfall — {input_wav_file_path} — {r.channel_name}")
            target = None
            if output_basename is not None:
                base = Path(output_basename)
                target = base.with_name(f"{base.stem}_waterfall_{r.channel_name}.png").with_suffix(".png")
            finalize_and_show_or_save(figure, output_path=target, show_interactive=show_interactive)
    return results


def summarise_waterfall_results_text(results: List[ChannelWaterfallResult]) -> str:
    """`[ch] start_sample= dur= slices= f_bins=` per channel (reference waterfall.py:615-623)."""
    return "\n".join(
        f"[{r.channel_name}] start_sample={r.analysis_start_sample_index}  "
        f"dur={float(r.analysis_length_samples) / float(r.sample_rate_hz):.3f}s  "
        f"slices={int(r.slice_times_seconds.size)}  f_bins={int(r.frequency_hz.size)}"
        for r in results
    )
