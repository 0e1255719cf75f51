#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference imports matplotlib at module level (analyse/decay.py:31,
analyse/waterfall.py:30-32) and matplotlib is not installed, so dummy modules
are registered first.  Only the plotting functions become unusable; every
`analyse_*_for_channel` runs the reference's own numpy code.

Inputs are stored as int16 PCM (the bundle wire format, recorder.hpp:49-53)
and converted with the reference's own rule x = pcm / 32768 (analyse/io.py:59),
so the fixtures carry exactly the float32 samples the reference analysed.
"""

from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path("/root/reference")


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Dummy(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return _Dummy("call")


def import_reference():
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(mod, _Dummy(mod))
    sys.path.insert(0, str(REFERENCE_ROOT))
    import analyse.decay as rdecay
    import analyse.rt60bands as rbands
    import analyse.waterfall as rwf
    import analyse.deconvolve as rdec

    return rdecay, rbands, rwf, rdec


def synth_pcm16(n: int, sr: int, rt60: float, onset: int, seed: int, floor_db: float = -90.0) -> np.ndarray:
    """Decaying gaussian noise with `onset` samples of leading silence."""
    rng = np.random.default_rng(seed)
    x = np.zeros(n, dtype=np.float64)
    t = np.arange(n - onset, dtype=np.float64) / sr
    x[onset:] = rng.standard_normal(n - onset) * 10.0 ** (-3.0 * t / rt60)
    x += 10.0 ** (floor_db / 20.0) * rng.standard_normal(n)
    x *= 0.95 / np.max(np.abs(x))
    return np.round(x * 32767.0).astype(np.int16)


def fits_to_rows(fits: dict) -> np.ndarray:
    """(3, 7) float64: rows EDT, T20, T30; cols t_start,t_end,slope,intercept,r2,rt60,present."""
    out = np.full((3, 7), np.nan)
    for r, name in enumerate(("EDT", "T20", "T30")):
        f = fits.get(name)
        if f is None:
            out[r, 6] = 0.0
            continue
        out[r] = [f.start_time_seconds, f.end_time_seconds, f.slope_db_per_second, f.intercept_db, f.r_squared,
                  f.rt60_seconds, 1.0]
    return out


def none_to_nan(v):
    return np.nan if v is None else float(v)


CASES = [
    # name, n, rt60, onset, seed      (lengths: 2^a 3^b 5^c, a 13-smooth one, an odd one, a large-prime one)
    ("smooth_48000", 48000, 0.30, 12000, 101),
    ("smooth_60000", 60000, 0.45, 14400, 102),
    ("radix13_62400", 62400, 0.35, 13000, 103),
    ("prime_factor_50022", 50022, 0.30, 12500, 104),  # 2 * 3^2 * 7 * 397 -> Bluestein on N/2
    ("odd_45001", 45001, 0.25, 12001, 105),
    ("onset_at_zero_48000", 48000, 0.30, 0, 106),
]


def main():
    rdecay, rbands, rwf, rdec = import_reference()
    sr = 48000
    manifest = {}
    for name, n, rt60, onset, seed in CASES:
        pcm = synth_pcm16(n, sr, rt60, onset, seed)
        x = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
        store = {"pcm": pcm, "sr": np.int32(sr)}

        # decay, EDT on (analyse/decay.py:268)
        d = rdecay.analyse_decay_for_channel(x, sr, "mono", rdecay.DecayAnalysisSettings(compute_edt=True))
        store["decay_start"] = np.int64(d.analysis_start_sample_index)
        store["decay_edc_db"] = d.edc_db
        store["decay_early10"] = np.float64(none_to_nan(d.early_decay_10db_time_seconds))
        store["decay_fits"] = fits_to_rows(d.fits)

        # decay with ignore_leading and smoothing variants (metrics only)
        d2 = rdecay.analyse_decay_for_channel(
            x, sr, "mono", rdecay.DecayAnalysisSettings(compute_edt=True, ignore_leading_seconds=0.01))
        store["decay_ign_start"] = np.int64(d2.analysis_start_sample_index)
        store["decay_ign_fits"] = fits_to_rows(d2.fits)
        store["decay_ign_early10"] = np.float64(none_to_nan(d2.early_decay_10db_time_seconds))
        d3 = rdecay.analyse_decay_for_channel(
            x, sr, "mono", rdecay.DecayAnalysisSettings(compute_edt=True, edc_smoothing_window_samples=31))
        store["decay_smooth_fits"] = fits_to_rows(d3.fits)
        store["decay_smooth_edc_db_dec16"] = d3.edc_db[::16].copy()
        d4 = rdecay.analyse_decay_for_channel(
            x, sr, "mono", rdecay.DecayAnalysisSettings(compute_edt=True, trim_to_peak=False))
        store["decay_notrim_fits"] = fits_to_rows(d4.fits)

        # rt60bands in all three modes with T20 + EDT (analyse/rt60bands.py:324)
        for mode in ("three", "octave", "third"):
            s = rbands.Rt60BandsAnalysisSettings(band_mode=mode, include_t20=True, include_edt=True)
            r = rbands.analyse_rt60_bands_for_channel(x, sr, "mono", s)
            tab = np.array(
                [[none_to_nan(r.band_metrics_by_name[b.name].rt60_t30_seconds),
                  none_to_nan(r.band_metrics_by_name[b.name].rt60_t20_seconds),
                  none_to_nan(r.band_metrics_by_name[b.name].edt_seconds)] for b in r.band_definitions])
            store[f"bands_{mode}_metrics"] = tab
            store[f"bands_{mode}_names"] = np.array([b.name for b in r.band_definitions])
            store[f"bands_{mode}_edges"] = np.array(
                [[none_to_nan(b.low_edge_hz), none_to_nan(b.high_edge_hz), b.centre_hz] for b in r.band_definitions])

        # one band signal + its EDC for a mid band, for direct filter parity (rt60bands.py:170, decay.py:115)
        s = rbands.Rt60BandsAnalysisSettings(band_mode="third")
        freqs = np.fft.rfftfreq(n, d=1.0 / sr).astype(np.float32)
        for probe in ("1000Hz", "63Hz" if False else "62Hz"):
            band = [b for b in rbands._build_band_definitions(s, sr) if b.name == probe][0]
            mask = rbands._make_bandpass_mask(freqs, band.low_edge_hz, band.high_edge_hz, s.transition_width_octaves,
                                              0.5 * sr)
            y = rbands._apply_fft_mask(x, mask)
            store[f"band_{probe}_signal_dec8"] = y[::8].copy()
            store[f"band_{probe}_mask_nonzero"] = np.array([int(np.argmax(mask > 0)),
                                                           int(mask.size - np.argmax(mask[::-1] > 0))])

        # waterfall default + slice_max variant (analyse/waterfall.py:349)
        if n - d.analysis_start_sample_index >= 4096:
            w = rwf.analyse_waterfall_for_channel(x, sr, "mono", rwf.WaterfallAnalysisSettings())
            store["wf_start"] = np.int64(w.analysis_start_sample_index)
            store["wf_length"] = np.int64(w.analysis_length_samples)
            store["wf_times"] = w.slice_times_seconds
            store["wf_freq"] = w.frequency_hz
            store["wf_rel_db"] = w.slice_magnitude_rel_db
            w2 = rwf.analyse_waterfall_for_channel(
                x, sr, "mono", rwf.WaterfallAnalysisSettings(db_reference="slice_max", num_slices=7,
                                                             slice_mode="uniform_frames", dynamic_range_db=60.0))
            store["wf2_times"] = w2.slice_times_seconds
            store["wf2_rel_db"] = w2.slice_magnitude_rel_db
            w3 = rwf.analyse_waterfall_for_channel(
                x, sr, "mono", rwf.WaterfallAnalysisSettings(smoothing_log_bins=9, slice_mode="uniform_time",
                                                             slice_spacing_seconds=0.08))
            store["wf3_times"] = w3.slice_times_seconds
            store["wf3_rel_db"] = w3.slice_magnitude_rel_db

        np.savez_compressed(HERE / f"{name}.npz", **store)
        manifest[name] = {"n": n, "rt60": rt60, "onset": onset, "seed": seed}
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in store.items() if hasattr(v, "shape")})

    # analytic exponential: exact linear EDC
    ana = {}
    for rt in (0.3, 1.2):
        n = int(sr * 2.0)
        x = (10.0 ** (-3.0 * np.arange(n) / (sr * rt))).astype(np.float32)
        d = rdecay.analyse_decay_for_channel(x, sr, "mono", rdecay.DecayAnalysisSettings(compute_edt=True))
        ana[f"rt{rt}_fits"] = fits_to_rows(d.fits)
        ana[f"rt{rt}_early10"] = np.float64(none_to_nan(d.early_decay_10db_time_seconds))
        ana[f"rt{rt}_edc_last"] = np.float32(d.edc_db[-1])
    np.savez_compressed(HERE / "analytic.npz", **ana)

    # deconvolution fixture (analyse/deconvolve.py:124): short sweep, 2-channel capture
    rng = np.random.default_rng(7)
    n_sw = 6000
    tt = np.arange(n_sw) / sr
    sweep = (0.5 * np.sin(2 * np.pi * (200.0 * tt + 0.5 * (8000.0 - 200.0) / tt[-1] * tt * tt))).astype(np.float32)
    ir = (rng.standard_normal((1500, 2)) * np.exp(-np.arange(1500) / 200.0)[:, None]).astype(np.float32)
    rec = np.stack([np.convolve(sweep, ir[:, c])[:9000] for c in range(2)], axis=1)
    rec = (0.9 * rec / np.abs(rec).max()).astype(np.float32)
    h = rdec.deconvolve_impulse_response(rec, sweep, sr, rdec.DeconvolveSettings())
    np.savez_compressed(HERE / "deconvolve.npz", sweep=sweep, recorded=rec, ir=h)

    (HERE / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print("done")


if __name__ == "__main__":
    main()
