"""numpy restatement of the reference algorithms on the hot path (oracle).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py for who may import this.

Every function cites the reference lines it restates (paths relative to the
kianmcevoy/audio_analysis tree).  The arithmetic dtypes follow the reference
exactly (float64 FFT / cumsum / lstsq, float32 masks, float32 EDC and time
vectors, NEP-50 weak python scalars), because the fits run on the float32
rounded curves.  Results are plain tuples / dicts, not the reference's
dataclasses: the product package owns the dataclass mirror.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------
# parameter bundles (defaults = reference dataclass defaults)
# ----------------------------------------------------------------------------


@dataclass(frozen=True)
class DecayParams:
    """analyse/decay.py:44-72 (DecayAnalysisSettings), analysis fields only."""

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
class BandParams:
    """analyse/rt60bands.py:44-69 (Rt60BandsAnalysisSettings)."""

    band_mode: str = "three"
    low_upper_hz: float = 250.0
    mid_center_hz: float = 1000.0
    mid_width_octaves: float = 2.0
    high_lower_hz: float = 4000.0
    f_min_hz: float = 31.5
    f_max_hz: float = 16000.0
    transition_width_octaves: float = 1.0 / 6.0
    include_t20: bool = False
    include_edt: bool = False
    decay: DecayParams = field(default_factory=DecayParams)


@dataclass(frozen=True)
class CsdParams:
    """analyse/waterfall.py:43-80 (WaterfallAnalysisSettings), analysis fields."""

    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    analysis_duration_seconds: Optional[float] = None
    n_fft: int = 4096
    hop_length: int = 512
    use_hann_window: bool = True
    f_min_hz: float = 20.0
    f_max_hz: float = 20000.0
    slice_mode: str = "auto"
    num_slices: int = 18
    slice_spacing_seconds: float = 0.05
    start_time_seconds: float = 0.0
    end_time_seconds: Optional[float] = None
    db_reference: str = "global_max"
    smoothing_log_bins: int = 0
    log_bins_per_octave: int = 96
    dynamic_range_db: float = 80.0
    floor_db: float = -120.0


# ----------------------------------------------------------------------------
# decay: Schroeder EDC, crossings, line fits
# ----------------------------------------------------------------------------


def schroeder_edc_db(x: np.ndarray, sr: int, p: DecayParams):
    """analyse/decay.py:115-170.  Returns (time_s f32, edc_db f32, start_index)."""
    if x.ndim != 1:
        raise ValueError("compute_schroeder_edc_db expects a 1D mono array.")
    seg = x.astype(np.float64, copy=False)
    start = 0
    if p.trim_to_peak:  # decay.py:135-138
        start = int(np.argmax(np.abs(seg)))
        seg = seg[start:]
    if p.ignore_leading_seconds > 0.0:  # decay.py:140-144
        skip = int(round(p.ignore_leading_seconds * float(sr)))
        skip = max(0, min(skip, len(seg)))
        start += skip
        seg = seg[skip:]
    if seg.size < 4:  # decay.py:146
        raise ValueError("Not enough samples after trimming/ignoring to compute EDC.")

    tail_energy = np.cumsum((seg * seg)[::-1])[::-1]  # decay.py:150-151
    tail_energy = np.maximum(tail_energy, float(p.edc_epsilon))  # :154
    tail_energy = tail_energy / tail_energy[0]  # :157
    curve = 10.0 * np.log10(tail_energy)  # :158
    w = int(p.edc_smoothing_window_samples or 0)
    if w > 1:  # :161-164, box filter in the dB domain, zero padded ("same")
        curve = np.convolve(curve, np.full(w, 1.0 / float(w)), mode="same")
    curve = np.maximum(curve, float(p.edc_floor_db)).astype(np.float32)  # :167
    t = (np.arange(curve.size, dtype=np.float32) / float(sr)).astype(np.float32)  # :169
    return t, curve, start


def crossing_time(t: np.ndarray, curve: np.ndarray, target_db: float) -> Optional[float]:
    """analyse/decay.py:173-199.  First sample with curve <= target, interpolated."""
    hit = curve <= target_db  # float32 compare against a weak python scalar
    if not hit.any():
        return None
    i = int(np.argmax(hit))
    if i == 0:
        return float(t[0])
    ta, tb = float(t[i - 1]), float(t[i])
    ya, yb = float(curve[i - 1]), float(curve[i])
    if yb == ya:
        return tb
    f = (target_db - ya) / (yb - ya)
    f = float(np.clip(f, 0.0, 1.0))
    return ta + f * (tb - ta)


def line_fit(t: np.ndarray, curve: np.ndarray, range_db, lower_limit_db: float) -> Optional[dict]:
    """analyse/decay.py:202-260.  None when the range is unavailable."""
    hi, lo = float(range_db[0]), float(range_db[1])
    if lo > hi:
        raise ValueError("range_db should be (higher_db, lower_db), e.g. (-5, -25).")
    lo_eff = max(lo, float(lower_limit_db))  # :219
    t_a = crossing_time(t, curve, hi)
    t_b = crossing_time(t, curve, lo_eff)
    if t_a is None or t_b is None or t_b <= t_a:  # :224
        return None
    sel = (t >= t_a) & (t <= t_b)  # :227 (float32 compare, weak scalars)
    if int(sel.sum()) < 8:  # :228
        return None
    tt = t[sel].astype(np.float64)
    yy = curve[sel].astype(np.float64)
    design = np.column_stack([tt, np.ones_like(tt)])
    sol, _, _, _ = np.linalg.lstsq(design, yy, rcond=None)  # :235-236
    m, b = float(sol[0]), float(sol[1])
    if m >= 0.0:  # :241
        return None
    resid = yy - (m * tt + b)
    ss_res = float(np.sum(resid ** 2))
    ss_tot = float(np.sum((yy - np.mean(yy)) ** 2))
    r2 = 1.0 - ss_res / ss_tot if ss_tot > 0.0 else 0.0  # :244-247
    return {
        "range_db": (hi, lo),
        "t_start": float(t_a),
        "t_end": float(t_b),
        "slope": m,
        "intercept": b,
        "r2": float(r2),
        "rt60": float(-60.0 / m),  # :249
        "n_points": int(sel.sum()),
    }


def decay_metrics(x: np.ndarray, sr: int, p: DecayParams) -> dict:
    """analyse/decay.py:268-329 (analyse_decay_for_channel)."""
    t, curve, start = schroeder_edc_db(x, sr, p)
    t0 = crossing_time(t, curve, 0.0)  # :280
    t10 = crossing_time(t, curve, -10.0)  # :281
    early10 = float(t10 - t0) if (t0 is not None and t10 is not None and t10 >= t0) else None
    fits: Dict[str, dict] = {}
    order = []
    if p.compute_edt:
        order.append(("EDT", p.edt_range_db))
    order += [("T20", p.t20_range_db), ("T30", p.t30_range_db)]
    for name, rng in order:
        f = line_fit(t, curve, rng, p.fit_lower_limit_db)
        if f is not None:
            fits[name] = f
    return {"start": start, "time_s": t, "edc_db": curve, "early10": early10, "fits": fits}


# ----------------------------------------------------------------------------
# rt60bands: band table, raised-cosine masks, circular FFT filter
# ----------------------------------------------------------------------------


def band_table(p: BandParams, sr: int) -> List[dict]:
    """analyse/rt60bands.py:183-264.  dicts: name, centre, kind, low, high."""
    nyq = 0.5 * float(sr)
    mode = str(p.band_mode).lower()
    if mode == "three":  # :183-205
        lo_up = float(np.clip(p.low_upper_hz, 20.0, nyq))
        mid_c = float(np.clip(p.mid_center_hz, 20.0, nyq))
        width = float(max(0.1, p.mid_width_octaves))
        spread = float(2.0 ** float(0.5 * width))
        mid_lo = float(np.clip(mid_c / spread, 20.0, nyq))
        mid_hi = float(np.clip(mid_c * spread, 20.0, nyq))
        hi_lo = float(np.clip(p.high_lower_hz, 20.0, nyq))
        return [
            dict(name="Low", centre=float(np.sqrt(20.0 * lo_up)), kind="lowpass", low=None, high=lo_up),
            dict(name="Mid", centre=mid_c, kind="bandpass", low=mid_lo, high=mid_hi),
            dict(name="High", centre=float(np.sqrt(max(20.0, hi_lo) * nyq)), kind="highpass", low=hi_lo, high=None),
        ]
    if mode in ("octave", "third"):  # :208-253
        per_oct = 1.0 if mode == "octave" else 3.0
        f_lo = float(max(20.0, min(p.f_min_hz, nyq)))
        f_hi = float(max(f_lo, min(p.f_max_hz, nyq)))
        step = 2.0 ** (1.0 / per_oct)
        half = 2.0 ** (1.0 / (2.0 * per_oct))
        k0 = int(np.floor(np.log(f_lo / 1000.0) / np.log(step)))
        k1 = int(np.ceil(np.log(f_hi / 1000.0) / np.log(step)))
        out = []
        for k in range(k0, k1 + 1):
            fc = 1000.0 * (step ** float(k))
            if fc < f_lo or fc > f_hi:
                continue
            a = float(np.clip(fc / half, 20.0, nyq))
            b = float(np.clip(fc * half, 20.0, nyq))
            if b <= a:
                continue
            out.append(dict(name=f"{int(round(fc))}Hz", centre=float(fc), kind="bandpass", low=a, high=b))
        out.sort(key=lambda d: d["centre"])
        return out
    raise ValueError(f"Unknown band_mode: {p.band_mode}")


def _ramp(f: np.ndarray, a: float, b: float) -> np.ndarray:
    """analyse/rt60bands.py:116-124: half-cosine 0 -> 1 between a and b (float32)."""
    if b <= a:
        return (f >= b).astype(np.float32)
    u = np.clip((f - a) / (b - a), 0.0, 1.0)
    return (0.5 - 0.5 * np.cos(np.pi * u)).astype(np.float32)


def _lowpass(f, pass_hz, tw_oct, nyq):
    """analyse/rt60bands.py:127-137."""
    pass_hz = float(np.clip(pass_hz, 1.0, nyq))
    stop = float(min(nyq, pass_hz * float(2.0 ** float(tw_oct))))
    if stop <= pass_hz:
        stop = min(nyq, pass_hz + 1.0)
    m = 1.0 - _ramp(f, pass_hz, stop)
    m[f <= pass_hz] = 1.0
    m[f >= stop] = 0.0
    return m.astype(np.float32)


def _highpass(f, pass_hz, tw_oct, nyq):
    """analyse/rt60bands.py:140-150."""
    pass_hz = float(np.clip(pass_hz, 1.0, nyq))
    stop = float(max(1.0, pass_hz / float(2.0 ** float(tw_oct))))
    if pass_hz <= stop:
        stop = max(1.0, pass_hz - 1.0)
    m = _ramp(f, stop, pass_hz)
    m[f <= stop] = 0.0
    m[f >= pass_hz] = 1.0
    return m.astype(np.float32)


def band_mask(freqs_f32: np.ndarray, band: dict, tw_oct: float, nyq: float) -> np.ndarray:
    """analyse/rt60bands.py:153-167 and the kind dispatch at :363-389."""
    kind = band["kind"]
    if kind == "lowpass":
        return _lowpass(freqs_f32, band["high"], tw_oct, nyq)
    if kind == "highpass":
        return _highpass(freqs_f32, band["low"], tw_oct, nyq)
    if kind == "bandpass":
        a = float(np.clip(band["low"], 1.0, nyq))
        b = float(np.clip(band["high"], 1.0, nyq))
        if b <= a:
            return np.zeros_like(freqs_f32, dtype=np.float32)
        return (_highpass(freqs_f32, a, tw_oct, nyq) * _lowpass(freqs_f32, b, tw_oct, nyq)).astype(np.float32)
    raise ValueError(f"Unknown band kind: {kind}")


def band_filter(x_f32: np.ndarray, mask_f32: np.ndarray, spectrum: Optional[np.ndarray] = None) -> np.ndarray:
    """analyse/rt60bands.py:170-175: circular zero-phase filter over the WHOLE buffer.

    `spectrum` lets the caller reuse rfft(x); the reference recomputes it for
    every band, which changes nothing numerically.
    """
    n = int(x_f32.size)
    if spectrum is None:
        spectrum = np.fft.rfft(x_f32.astype(np.float64, copy=False))
    return np.fft.irfft(spectrum * mask_f32.astype(np.float64), n=n).astype(np.float32)


def band_metrics(x: np.ndarray, sr: int, p: BandParams) -> dict:
    """analyse/rt60bands.py:324-413 (analyse_rt60_bands_for_channel).

    Returns {"bands": [band dicts], "metrics": {name: (t30, t20, edt)}, "start": idx}.
    """
    full = x.astype(np.float32, copy=False)
    dp = p.decay
    peak = int(np.argmax(np.abs(full))) if dp.trim_to_peak else 0  # :333-335
    skip = 0
    if dp.ignore_leading_seconds > 0.0:  # :337-340
        skip = int(round(dp.ignore_leading_seconds * float(sr)))
        skip = max(0, min(skip, full.size))
    start = min(full.size, peak + skip)  # :342
    n = int(full.size)
    if n < 8:  # :346
        raise ValueError("Not enough samples for rt60bands analysis.")
    nyq = 0.5 * float(sr)
    freqs = np.fft.rfftfreq(n, d=1.0 / float(sr)).astype(np.float32)  # :350
    bands = band_table(p, sr)
    # band EDC starts at the shared index, so its own trimming is switched off (:356-360)
    bdp = DecayParams(
        trim_to_peak=False,
        ignore_leading_seconds=0.0,
        edc_floor_db=dp.edc_floor_db,
        edc_epsilon=dp.edc_epsilon,
        fit_lower_limit_db=dp.fit_lower_limit_db,
        t20_range_db=dp.t20_range_db,
        t30_range_db=dp.t30_range_db,
        compute_edt=dp.compute_edt,
        edt_range_db=dp.edt_range_db,
        edc_smoothing_window_samples=dp.edc_smoothing_window_samples,
    )
    spec = np.fft.rfft(full.astype(np.float64, copy=False))
    out: Dict[str, Tuple[Optional[float], Optional[float], Optional[float]]] = {}
    for b in bands:
        y = band_filter(full, band_mask(freqs, b, p.transition_width_octaves, nyq), spec)[start:]  # :392-393
        if y.size < 8:  # :395-397
            out[b["name"]] = (None, None, None)
            continue
        t, curve, _ = schroeder_edc_db(y, sr, bdp)  # :279
        pick = lambda f: None if f is None else f["rt60"]
        t30 = pick(line_fit(t, curve, bdp.t30_range_db, bdp.fit_lower_limit_db))
        t20 = pick(line_fit(t, curve, bdp.t20_range_db, bdp.fit_lower_limit_db)) if p.include_t20 else None
        edt = pick(line_fit(t, curve, bdp.edt_range_db, bdp.fit_lower_limit_db)) if p.include_edt else None
        out[b["name"]] = (t30, t20, edt)
    return {"bands": bands, "metrics": out, "start": start}


# ----------------------------------------------------------------------------
# waterfall / CSD
# ----------------------------------------------------------------------------


def stft_db(x_f32: np.ndarray, sr: int, n_fft: int, hop: int, hann: bool, floor_db: float,
            frames: Optional[Sequence[int]] = None):
    """analyse/waterfall.py:188-230.  Returns (frame_times f32[T], freq f32[F], mag_db f32[F, T']).

    `frames` restricts the columns computed (the reference computes all T and
    later uses <= num_slices of them); T' = len(frames) or T.
    """
    if x_f32.ndim != 1:
        raise ValueError("_compute_stft_magnitude_db expects a 1D mono array.")
    if n_fft <= 0 or hop <= 0:
        raise ValueError("n_fft and hop_length must be positive.")
    if x_f32.size < n_fft:
        raise ValueError("Not enough samples for STFT (need at least n_fft).")
    x = x_f32.astype(np.float64, copy=False)
    n_frames = 1 + (x.size - n_fft) // hop
    win = np.hanning(n_fft).astype(np.float64) if hann else np.ones(n_fft)
    freq = np.fft.rfftfreq(n_fft, d=1.0 / float(sr)).astype(np.float32)
    lin_floor = 10.0 ** (float(floor_db) / 20.0)
    cols = range(n_frames) if frames is None else [int(i) for i in frames]
    out = np.empty((freq.size, len(cols)), dtype=np.float32)
    for c, fi in enumerate(cols):
        seg = x[fi * hop: fi * hop + n_fft] * win
        mag = np.maximum(np.abs(np.fft.rfft(seg)), lin_floor)
        out[:, c] = (20.0 * np.log10(mag)).astype(np.float32)
    times = (np.arange(n_frames, dtype=np.float32) * float(hop) / float(sr)).astype(np.float32)
    return times, freq, out


def pick_slices(frame_times: np.ndarray, p: CsdParams) -> np.ndarray:
    """analyse/waterfall.py:233-286."""
    if frame_times.size == 0:
        return np.zeros((0,), dtype=np.int32)
    t_a = float(max(0.0, p.start_time_seconds))
    t_b = float(p.end_time_seconds) if p.end_time_seconds is not None else float(frame_times[-1])
    if t_b <= t_a:
        t_b = float(frame_times[-1])
    ok = (frame_times >= t_a) & (frame_times <= t_b)
    if not ok.any():
        return np.zeros((0,), dtype=np.int32)
    i_lo = int(np.argmax(ok))
    i_hi = int(np.max(np.nonzero(ok)))
    mode = str(p.slice_mode).lower()
    if mode == "uniform_frames":
        return np.unique(np.linspace(i_lo, i_hi, int(max(1, p.num_slices)), dtype=np.int32))
    if mode == "uniform_time":
        step = float(max(1e-4, p.slice_spacing_seconds))
        wanted = np.arange(t_a, t_b + 1e-9, step, dtype=np.float64)
    else:  # auto
        wanted = np.linspace(t_a, t_b, int(max(2, p.num_slices)), dtype=np.float64)
    chosen = []
    for tt in wanted:
        j = int(np.argmin(np.abs(frame_times - float(tt))))
        if i_lo <= j <= i_hi:
            chosen.append(j)
    if mode == "uniform_time" and not chosen:
        chosen = [i_lo, i_hi]
    return np.unique(np.array(chosen, dtype=np.int32))


def _box(x: np.ndarray, w: int) -> np.ndarray:
    """analyse/waterfall.py:132-137."""
    if w <= 1:
        return x.astype(np.float32, copy=False)
    return np.convolve(x.astype(np.float64), np.full(int(w), 1.0 / float(int(w))), mode="same").astype(np.float32)


def _smooth_logf(freq: np.ndarray, mag_db: np.ndarray, f_lo: float, f_hi: float, w: int, per_oct: int) -> np.ndarray:
    """analyse/waterfall.py:140-185."""
    if w <= 1:
        return mag_db.astype(np.float32, copy=False)
    f = freq.astype(np.float64, copy=False)
    m = mag_db.astype(np.float64, copy=False)
    lo = float(max(1.0, f_lo))
    hi = float(max(lo, f_hi))
    sel = (f >= lo) & (f <= hi)
    if not sel.any():
        return mag_db.astype(np.float32, copy=False)
    fs, ms = f[sel], m[sel]
    l0, l1 = float(np.log2(fs[0])), float(np.log2(fs[-1]))
    bpo = int(max(16, per_oct))
    count = int(max(8, np.ceil((l1 - l0) * bpo))) + 1
    grid = 2.0 ** np.linspace(l0, l1, count, dtype=np.float64)
    g = np.interp(grid, fs, ms)
    g = _box(g.astype(np.float32), int(w)).astype(np.float64)
    back = np.interp(fs, grid, g)
    out = mag_db.astype(np.float32, copy=True)
    out[sel] = back.astype(np.float32)
    return out


def csd_slices(freq: np.ndarray, mag_db_cols: np.ndarray, p: CsdParams):
    """analyse/waterfall.py:289-341.  mag_db_cols is (F_all, S): columns already gathered."""
    nyq = float(freq[-1]) if freq.size else 0.0
    f_lo = float(np.clip(p.f_min_hz, 1.0, nyq))
    f_hi = float(np.clip(p.f_max_hz, f_lo, nyq))
    keep = (freq >= f_lo) & (freq <= f_hi)
    if not keep.any():
        raise ValueError("Waterfall frequency selection is empty (check f_min_hz/f_max_hz).")
    f_sel = freq[keep].astype(np.float32)
    rows = mag_db_cols[keep].T.astype(np.float32)  # (S, F)
    if p.smoothing_log_bins and int(p.smoothing_log_bins) > 1:
        rows = np.stack([
            _smooth_logf(f_sel, r, f_lo, f_hi, int(p.smoothing_log_bins), int(p.log_bins_per_octave)) for r in rows
        ], axis=0).astype(np.float32)
    if str(p.db_reference).lower() == "slice_max":
        rel = rows - np.max(rows, axis=1, keepdims=True)
    else:
        rel = rows - float(np.max(rows))
    dyn = float(max(10.0, p.dynamic_range_db))
    return f_sel, np.clip(rel, -dyn, 0.0).astype(np.float32)


def waterfall(x: np.ndarray, sr: int, p: CsdParams, only_selected_frames: bool = True) -> dict:
    """analyse/waterfall.py:349-410 (analyse_waterfall_for_channel).

    only_selected_frames=True computes just the <= num_slices STFT columns that
    reach the result (identical output; the reference computes every frame).
    """
    seg = x.astype(np.float64, copy=False)
    start = 0
    if p.trim_to_peak:
        start = int(np.argmax(np.abs(seg)))
        seg = seg[start:]
    if p.ignore_leading_seconds > 0.0:
        skip = int(round(float(p.ignore_leading_seconds) * float(sr)))
        skip = max(0, min(skip, seg.size))
        start += skip
        seg = seg[skip:]
    if p.analysis_duration_seconds is not None:
        keep = int(round(float(p.analysis_duration_seconds) * float(sr)))
        seg = seg[: max(0, min(keep, seg.size))]
    seg32 = seg.astype(np.float32)
    if seg32.size < p.n_fft:
        raise ValueError("Not enough samples after trimming/selection for waterfall (need at least n_fft).")
    n_frames = 1 + (seg32.size - int(p.n_fft)) // int(p.hop_length)
    times = (np.arange(n_frames, dtype=np.float32) * float(p.hop_length) / float(sr)).astype(np.float32)
    idx = pick_slices(times, p)
    if idx.size < 2:
        raise ValueError("Not enough slices selected for waterfall (increase duration or num_slices).")
    _, freq, cols = stft_db(seg32, sr, int(p.n_fft), int(p.hop_length), bool(p.use_hann_window), float(p.floor_db),
                            frames=idx if only_selected_frames else None)
    if not only_selected_frames:
        cols = cols[:, idx]
    f_sel, rel = csd_slices(freq, cols, p)
    return {
        "start": int(start),
        "length": int(seg32.size),
        "frame_indices": idx,
        "slice_times": times[idx].astype(np.float32),
        "freq": f_sel,
        "rel_db": rel,
    }


# ----------------------------------------------------------------------------
# spectrogram
# ----------------------------------------------------------------------------


@dataclass
class SpectrogramParams:
    """analyse/spectrogram.py:37-64 (the fields that reach the numbers)."""

    trim_to_peak: bool = True
    ignore_leading_seconds: float = 0.0
    analysis_duration_seconds: Optional[float] = None
    n_fft: int = 4096
    hop_length: int = 512
    use_hann_window: bool = True
    floor_db: float = -120.0


def spectrogram(x: np.ndarray, sr: int, p: SpectrogramParams) -> dict:
    """analyse/spectrogram.py:163-216 (analyse_spectrogram_for_channel) on top of :107-160, which is the same
    STFT loop as analyse/waterfall.py:188-230 (`stft_db`) with every frame kept."""
    if x.ndim != 1:
        raise ValueError("analyse_spectrogram_for_channel expects a 1D mono array.")
    seg = x.astype(np.float64, copy=False)
    start = 0
    if p.trim_to_peak:                      # :176-179
        start = int(np.argmax(np.abs(seg)))
        seg = seg[start:]
    if p.ignore_leading_seconds > 0.0:      # :181-185
        skip = int(round(float(p.ignore_leading_seconds) * float(sr)))
        skip = max(0, min(skip, seg.size))
        start += skip
        seg = seg[skip:]
    if p.analysis_duration_seconds is not None:   # :187-190
        keep = int(round(float(p.analysis_duration_seconds) * float(sr)))
        seg = seg[: max(0, min(keep, seg.size))]
    seg32 = seg.astype(np.float32)          # :192
    if seg32.size < p.n_fft:
        raise ValueError("Not enough samples after trimming/selection for spectrogram (need at least n_fft).")
    times, freq, mag = stft_db(seg32, sr, int(p.n_fft), int(p.hop_length), bool(p.use_hann_window), float(p.floor_db))
    return {"start": int(start), "length": int(seg32.size), "times": times, "freq": freq, "mag_db": mag}


# ----------------------------------------------------------------------------
# modal cloud
# ----------------------------------------------------------------------------


@dataclass
class ModalParams:
    """analyse/modalcloud.py:45-84 (the fields that reach the numbers)."""

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
    metric: str = "t30"
    min_fit_points: int = 10
    min_peak_db_above_floor: float = 20.0


def modal_log_bins(f_min: float, f_max: float, per_octave: int, min_bins: int) -> np.ndarray:
    """analyse/modalcloud.py:168-173."""
    import math

    lo = float(max(1.0, f_min))
    hi = float(max(lo * 1.001, f_max))
    octs = float(math.log2(hi / lo))
    n = int(max(min_bins, math.ceil(octs * float(max(4, per_octave)))))
    return (lo * (2.0 ** (np.linspace(0.0, octs, n + 1, dtype=np.float64)))).astype(np.float32)


def modal_aggregate(freq: np.ndarray, mag_db: np.ndarray, edges_f32: np.ndarray):
    """analyse/modalcloud.py:176-207: mean of linear magnitudes per log bin, back to dB; NaN rows for empty bins."""
    edges = edges_f32.astype(np.float64)
    centres = np.sqrt(edges[:-1] * edges[1:]).astype(np.float32)
    lin = 10.0 ** (mag_db.astype(np.float64) / 20.0)
    out = np.full((centres.size, mag_db.shape[1]), np.nan, dtype=np.float32)
    for b in range(centres.size):
        sel = (freq >= float(edges[b])) & (freq < float(edges[b + 1]))
        if not np.any(sel):
            continue
        m = np.maximum(np.mean(lin[sel, :], axis=0), 1e-30)
        out[b, :] = (20.0 * np.log10(m)).astype(np.float32)
    return centres, out


def modal_fit(t: np.ndarray, curve: np.ndarray, range_db, lower_limit_db: float, min_points: int):
    """analyse/modalcloud.py:239-281.  (slope, intercept, r2, rt60) or None."""
    hi, lo = float(range_db[0]), float(range_db[1])
    if lo > hi:
        raise ValueError("range_db should be (higher_db, lower_db), e.g. (-5, -35).")
    t0 = crossing_time(t, curve, hi)
    t1 = crossing_time(t, curve, max(lo, float(lower_limit_db)))
    if t0 is None or t1 is None or t1 <= t0:
        return None
    sel = (t >= t0) & (t <= t1)
    if int(np.sum(sel)) < int(min_points):
        return None
    tt = t[sel].astype(np.float64)
    yy = curve[sel].astype(np.float64)
    coeffs, _, _, _ = np.linalg.lstsq(np.column_stack([tt, np.ones_like(tt)]), yy, rcond=None)
    slope, icpt = float(coeffs[0]), float(coeffs[1])
    if slope >= 0.0:
        return None
    pred = slope * tt + icpt
    ss_res = float(np.sum((yy - pred) ** 2))
    ss_tot = float(np.sum((yy - np.mean(yy)) ** 2))
    r2 = 1.0 - (ss_res / ss_tot) if ss_tot > 0.0 else 0.0
    return slope, icpt, r2, float(-60.0 / slope)


def modal_cloud(x: np.ndarray, sr: int, p: ModalParams) -> dict:
    """analyse/modalcloud.py:284-391 (analyse_modal_cloud_for_channel)."""
    seg = x.astype(np.float64, copy=False)
    start = 0
    if p.trim_to_peak:
        start = int(np.argmax(np.abs(seg)))
        seg = seg[start:]
    if p.ignore_leading_seconds > 0.0:
        skip = int(round(float(p.ignore_leading_seconds) * float(sr)))
        skip = max(0, min(skip, seg.size))
        start += skip
        seg = seg[skip:]
    if p.analysis_duration_seconds is not None:
        keep = int(round(float(p.analysis_duration_seconds) * float(sr)))
        seg = seg[: max(0, min(keep, seg.size))]
    seg32 = seg.astype(np.float32)
    if seg32.size < p.n_fft:
        raise ValueError("Not enough samples after trimming/selection for modal cloud (need at least n_fft).")
    times, freq, mag = stft_db(seg32, sr, int(p.n_fft), int(p.hop_length), bool(p.use_hann_window), float(p.floor_db))
    nyq = 0.5 * float(sr)
    f_lo = float(np.clip(p.f_min_hz, 1.0, nyq))
    f_hi = float(np.clip(p.f_max_hz, f_lo, nyq))
    keep_f = (freq >= f_lo) & (freq <= f_hi)                      # :324-327
    edges = modal_log_bins(f_lo, f_hi, int(p.log_bins_per_octave), int(p.min_bins))
    centres, curves = modal_aggregate(freq[keep_f], mag[keep_f, :], edges)
    metric = str(p.metric).lower()
    if metric == "t20":
        rng = p.t20_range_db
    elif metric == "edt":
        rng = p.edt_range_db
    else:
        metric, rng = "t30", p.t30_range_db
    points = []
    for b in range(centres.size):
        curve = curves[b, :]
        if not np.all(np.isfinite(curve)):
            continue
        peak = float(np.max(curve))
        if (peak - float(p.floor_db)) < float(p.min_peak_db_above_floor):
            continue
        rel = (curve - peak).astype(np.float32)
        fit = modal_fit(times, rel, rng, float(p.fit_lower_limit_db), int(p.min_fit_points))
        if fit is None:
            continue
        points.append((float(centres[b]), float(fit[3]), float(fit[2])))
    points.sort(key=lambda q: q[0])
    return {"start": int(start), "length": int(seg32.size), "metric": metric, "points": points}


# ----------------------------------------------------------------------------
# deconvolution (input producer for config 4)
# ----------------------------------------------------------------------------


def deconvolve(recorded_2d: np.ndarray, sweep_1d: np.ndarray, reg_rel: float = 1e-10, normalise_peak: bool = True,
               target_peak: float = 0.95, remove_dc: bool = True, full_fft: bool = False) -> np.ndarray:
    """analyse/deconvolve.py:124-193.  recorded_2d is float32 (N, C) in [-1, 1]."""
    rec = np.clip(np.asarray(recorded_2d, dtype=np.float32), -1.0, 1.0)
    if rec.ndim == 1:
        rec = rec.reshape(-1, 1)
    sweep = np.asarray(sweep_1d, dtype=np.float32)
    if rec.shape[0] < 8 or sweep.size < 8:
        raise ValueError("Recorded and sweep must both contain at least a few samples.")
    n_rec = int(rec.shape[0])
    n = max(n_rec, int(sweep.size))
    n_fft = 1 if n <= 1 else 1 << (int(n - 1).bit_length())
    X = np.fft.rfft(sweep.astype(np.float64), n=n_fft)
    pw = np.abs(X) ** 2
    eps = float(reg_rel) * max(1e-30, float(pw.max()))
    cols = []
    for c in range(rec.shape[1]):
        Y = np.fft.rfft(rec[:, c].astype(np.float64), n=n_fft)
        h = np.fft.irfft((Y * np.conj(X)) / (pw + eps), n=n_fft).astype(np.float32)
        if not full_fft:
            h = h[:n_rec]
        if remove_dc and h.size:
            h = (h - float(np.mean(h))).astype(np.float32)
        cols.append(h)
    out = np.stack(cols, axis=1).astype(np.float32)
    if normalise_peak:
        pk = float(np.max(np.abs(out))) if out.size else 0.0
        if pk > 0.0:
            out = (out * (float(target_peak) / pk)).astype(np.float32)
    return out
