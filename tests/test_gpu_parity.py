"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference fixtures.

Run on the B200 box with `pytest -m gpu`.  Tolerances are BASELINE.json's: EDC within 0.01 dB down to
-60 dB, T20/T30/RT60/EDT within 0.5 % or 1 ms, CSD magnitudes within 1e-4 relative (8.7e-4 dB).
"""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from _parity import (CSD_TOL_DB, FIT_ABS_TOL_S, FIT_REL_TOL, check_band_table, check_edc, check_fit_rows, note,
                     seconds_close)
from conftest import GOLDEN_CASES, load_golden, load_verb_case

pytestmark = pytest.mark.gpu

SR = 48000


def _fits_table(fits: dict) -> np.ndarray:
    out = np.full((3, 7), np.nan)
    for r, name in enumerate(("EDT", "T20", "T30")):
        f = fits.get(name)
        if f is None:
            out[r, 6] = 0.0
        elif isinstance(f, dict):
            out[r] = [f["t_start"], f["t_end"], f["slope"], f["intercept"], f["r2"], f["rt60"], 1.0]
        else:
            out[r] = [f.start_time_seconds, f.end_time_seconds, f.slope_db_per_second, f.intercept_db, f.r_squared,
                      f.rt60_seconds, 1.0]
    return out


def nat_lib():
    from audio_analysis_b200 import _native as nat
    return nat.load()


@pytest.fixture(scope="module")
def dec():
    from audio_analysis_b200 import decay
    return decay


# ------------------------------------------------------------------------------------ decay


def test_decay_matches_reference_fixtures(golden, dec):
    r = dec.analyse_decay_for_channel(golden["x"], SR, "mono", dec.DecayAnalysisSettings(compute_edt=True))
    assert r.analysis_start_sample_index == int(golden["decay_start"])
    assert r.edc_db.dtype == np.float32 and r.time_seconds.dtype == np.float32
    check_edc(golden["name"], r.edc_db, golden["decay_edc_db"])
    assert np.array_equal(r.time_seconds, (np.arange(r.edc_db.size, dtype=np.float32) / float(SR)))
    check_fit_rows(golden["name"], _fits_table(r.fits), golden["decay_fits"])
    want = float(golden["decay_early10"])
    assert (r.early_decay_10db_time_seconds is None) == np.isnan(want)
    if not np.isnan(want):
        note("early10_abs_err_s", abs(r.early_decay_10db_time_seconds - want))
        assert seconds_close(r.early_decay_10db_time_seconds, want)


def test_decay_variants(golden, dec):
    x = golden["x"]
    r = dec.analyse_decay_for_channel(x, SR, "m", dec.DecayAnalysisSettings(compute_edt=True, ignore_leading_seconds=0.01))
    assert r.analysis_start_sample_index == int(golden["decay_ign_start"])
    check_fit_rows(golden["name"] + "/ignore", _fits_table(r.fits), golden["decay_ign_fits"])
    r = dec.analyse_decay_for_channel(x, SR, "m", dec.DecayAnalysisSettings(compute_edt=True, trim_to_peak=False))
    assert r.analysis_start_sample_index == 0
    check_fit_rows(golden["name"] + "/notrim", _fits_table(r.fits), golden["decay_notrim_fits"])
    r = dec.analyse_decay_for_channel(x, SR, "m",
                                      dec.DecayAnalysisSettings(compute_edt=True, edc_smoothing_window_samples=31))
    check_fit_rows(golden["name"] + "/smooth", _fits_table(r.fits), golden["decay_smooth_fits"])
    got, want = r.edc_db[::16], golden["decay_smooth_edc_db_dec16"]
    sel = want >= -60.0
    assert np.max(np.abs(got[sel] - want[sel])) <= 0.01


def test_decay_ragged_batch_equals_single_rows(dec):
    cases = [load_golden(n) for n in GOLDEN_CASES]
    chans = [c["x"] for c in cases]
    s = dec.DecayAnalysisSettings(compute_edt=True)
    batch = dec.analyse_decay_batch(chans, SR, GOLDEN_CASES, s)
    for c, got in zip(cases, batch):
        assert got.analysis_start_sample_index == int(c["decay_start"])
        check_edc("ragged", got.edc_db, c["decay_edc_db"])
        check_fit_rows("ragged", _fits_table(got.fits), c["decay_fits"])
    # batch result == batch-size-1 result, bit for bit (the kernel is deterministic)
    single = dec.analyse_decay_for_channel(chans[1], SR, "x", s)
    assert np.array_equal(single.edc_db, batch[1].edc_db)
    # the float64 regression sums are reduced in a different order when the tile -> warp split changes
    for k, f in single.fits.items():
        g = batch[1].fits[k]
        assert f.start_time_seconds == g.start_time_seconds and f.end_time_seconds == g.end_time_seconds
        assert f.rt60_seconds == pytest.approx(g.rt60_seconds, rel=1e-12)
        assert f.r_squared == pytest.approx(g.r_squared, abs=1e-12)


def test_decay_analytic_exponential(dec):
    for rt in (0.3, 1.2, 2.5):
        n = int(SR * 3.0)
        x = (10.0 ** (-3.0 * np.arange(n) / (SR * rt))).astype(np.float32)
        r = dec.analyse_decay_for_channel(x, SR, "m", dec.DecayAnalysisSettings(compute_edt=True))
        assert r.analysis_start_sample_index == 0
        for name in ("EDT", "T20", "T30"):
            assert r.fits[name].rt60_seconds == pytest.approx(rt, rel=2e-4)
            assert r.fits[name].r_squared == pytest.approx(1.0, abs=1e-6)
        assert r.early_decay_10db_time_seconds == pytest.approx(rt / 6.0, rel=1e-3)
        want = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True))
        check_edc("analytic", r.edc_db, want["edc_db"])
        check_fit_rows("analytic", _fits_table(r.fits), _fits_table(want["fits"]))


def test_decay_long_rows_and_clusters(dec):
    """Row lengths that exercise every cluster size (1, 2, 4, 8 CTAs per row)."""
    from audio_analysis_b200.synth import synth_ir_numpy
    s = dec.DecayAnalysisSettings(compute_edt=True)
    for n, idx in ((9000, 1), (40000, 2), (100000, 3), (288000, 4), (624000, 5)):
        x = synth_ir_numpy(idx, n, SR)
        got = dec.analyse_decay_for_channel(x, SR, "m", s)
        want = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True))
        assert got.analysis_start_sample_index == want["start"]
        check_edc(f"n{n}", got.edc_db, want["edc_db"])
        check_fit_rows(f"n{n}", _fits_table(got.fits), _fits_table(want["fits"]))


def test_decay_edge_cases(dec):
    s = dec.DecayAnalysisSettings(compute_edt=True)
    with pytest.raises(ValueError, match="1D mono"):
        dec.compute_schroeder_edc_db(np.zeros((4, 2), np.float32), SR, s)
    with pytest.raises(ValueError, match="Not enough samples"):
        dec.compute_schroeder_edc_db(np.array([0, 0, 1.0], np.float32), SR, s)
    with pytest.raises(ValueError, match="Not enough samples"):  # peak is the last sample
        dec.analyse_decay_for_channel(np.array([0.0] * 50 + [1.0], np.float32), SR, "m", s)
    r = dec.analyse_decay_for_channel(np.zeros(1000, np.float32), SR, "m", s)  # all-zero: flat 0 dB, no fits
    assert np.all(r.edc_db == 0.0) and r.fits == {} and r.early_decay_10db_time_seconds is None
    t, e, start = dec.compute_schroeder_edc_db(load_golden("smooth_48000")["x"], SR, s)
    assert start == int(load_golden("smooth_48000")["decay_start"]) and e[0] == 0.0 and t[0] == 0.0
    with pytest.raises(ValueError, match="higher_db, lower_db"):
        dec.fit_decay_slope_over_db_range(t, e, (-25.0, -5.0), -80.0, "bad")


def test_fit_decay_slope_on_supplied_curve(dec):
    g = load_golden("smooth_60000")
    want = oracle.decay_metrics(g["x"], SR, oracle.DecayParams(compute_edt=True))
    t, e = want["time_s"], want["edc_db"]
    for name, rng in (("EDT", (0.0, -10.0)), ("T20", (-5.0, -25.0)), ("T30", (-5.0, -35.0))):
        fit = dec.fit_decay_slope_over_db_range(t, e, rng, -80.0, name)
        w = want["fits"][name]
        assert fit is not None and fit.name == name
        assert fit.start_time_seconds == pytest.approx(w["t_start"], abs=1e-9)
        assert fit.end_time_seconds == pytest.approx(w["t_end"], abs=1e-9)
        assert fit.rt60_seconds == pytest.approx(w["rt60"], rel=1e-6)
        assert fit.r_squared == pytest.approx(w["r2"], abs=1e-7)
    assert dec.fit_decay_slope_over_db_range(t, e, (-5.0, -300.0), -200.0, "none") is None


# ------------------------------------------------------------------------------------ rt60bands


@pytest.mark.parametrize("mode", ["three", "octave", "third"])
def test_rt60bands_match_reference_fixtures(golden, mode):
    from audio_analysis_b200 import rt60bands as rb
    s = rb.Rt60BandsAnalysisSettings(band_mode=mode, include_t20=True, include_edt=True)
    r = rb.analyse_rt60_bands_for_channel(golden["x"], SR, "mono", s)
    names = [b.name for b in r.band_definitions]
    assert names == [str(v) for v in golden[f"bands_{mode}_names"]]
    got = np.array([[np.nan if v is None else v for v in
                     (r.band_metrics_by_name[n].rt60_t30_seconds, r.band_metrics_by_name[n].rt60_t20_seconds,
                      r.band_metrics_by_name[n].edt_seconds)] for n in names])
    want = golden[f"bands_{mode}_metrics"]
    # onset-at-zero: the circular filters wrap their pre-ringing into the tail and the band EDCs are nearly flat
    #  - the ill-conditioned edge case.  Every cell is held to the same 0.5 % / 1 ms bar as the others
    # (round 1 allowed 10 % of them outside; none is).
    check_band_table(f"{golden['name']}/{mode}", got, want)


def test_band_signals_match_oracle():
    """The filter itself: irfft(rfft(x) * mask, n) on the whole buffer (rt60bands.py:170-175)."""
    from audio_analysis_b200 import engine, rt60bands as rb
    for case in ("smooth_48000", "radix13_62400", "prime_factor_50022", "odd_45001"):
        g = load_golden(case)
        x = g["x"]
        s = rb.Rt60BandsAnalysisSettings(band_mode="third")
        rows = engine.make_rows(x[None, :])
        _, bands, _, start, sig = rb.run_rt60_bands(rows, SR, s, want_band_signals=True)
        sig = sig.cpu().numpy()[0]
        freqs = np.fft.rfftfreq(x.size, d=1.0 / SR).astype(np.float32)
        tab = oracle.band_table(oracle.BandParams(band_mode="third"), SR)
        spec = np.fft.rfft(x.astype(np.float64))
        worst = 0.0
        for k, b in enumerate(tab):
            want = oracle.band_filter(x, oracle.band_mask(freqs, b, 1.0 / 6.0, 0.5 * SR), spec)
            scale = float(np.max(np.abs(want))) + 1e-30
            worst = max(worst, float(np.max(np.abs(sig[k] - want))) / scale)
        note(f"band_signal_max_err_rel_to_peak[{case}]", worst)
        assert worst < 2e-5, (case, worst)
        assert int(start.cpu()[0]) == int(g["decay_start"])
        for probe in ("1000Hz", "62Hz"):
            k = [b["name"] for b in tab].index(probe)
            ref = g[f"band_{probe}_signal_dec8"]
            assert np.max(np.abs(sig[k][::8] - ref)) <= 2e-5 * (np.max(np.abs(ref)) + 1e-30)


def test_rt60bands_batch_third_octave_matches_oracle():
    from audio_analysis_b200 import rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_numpy
    x = synth_batch_numpy(6, 96000, SR)
    s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    got = rb.analyse_rt60_bands_batch(x, SR, None, s)
    p = oracle.BandParams(band_mode="third", include_t20=True, include_edt=True)
    for i in range(x.shape[0]):
        w = oracle.band_metrics(x[i], SR, p)
        names = [b["name"] for b in w["bands"]]
        want = np.array([[np.nan if v is None else v for v in w["metrics"][n]] for n in names])
        g = got[i]
        have = np.array([[np.nan if v is None else v for v in
                          (g.band_metrics_by_name[n].rt60_t30_seconds, g.band_metrics_by_name[n].rt60_t20_seconds,
                           g.band_metrics_by_name[n].edt_seconds)] for n in names])
        check_band_table(f"synth{i}", have, want)


@pytest.mark.parametrize("n", [288000, 192000])
def test_rt60bands_specialised_fft_lengths_match_oracle(n):
    """Config-5 (288,000) and config-3 (192,000) row lengths run the compile-time specialised FFT passes."""
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_numpy
    x = synth_batch_numpy(2, n, SR, first_index=40)
    s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    rows = engine.make_rows(x)
    metrics, bands, plan, start, sig = rb.run_rt60_bands(rows, SR, s, want_band_signals=True)
    got = metrics.cpu().numpy()
    sig = sig.cpu().numpy()
    p = oracle.BandParams(band_mode="third", include_t20=True, include_edt=True)
    freqs = np.fft.rfftfreq(n, d=1.0 / SR).astype(np.float32)
    for i in range(2):
        w = oracle.band_metrics(x[i], SR, p)
        names = [b["name"] for b in w["bands"]]
        want = np.array([[np.nan if v is None else v for v in w["metrics"][nm]] for nm in names])
        check_band_table(f"n{n}/{i}", got[i], want)
        spec = np.fft.rfft(x[i].astype(np.float64))
        for k in (0, 12, 25):
            ref = oracle.band_filter(x[i], oracle.band_mask(freqs, w["bands"][k], 1.0 / 6.0, 0.5 * SR), spec)
            err = float(np.max(np.abs(sig[i, k] - ref))) / (float(np.max(np.abs(ref))) + 1e-30)
            note(f"band_signal_max_err_rel_to_peak[n{n}]", err)
            assert err < 2e-5


def test_rt60bands_smoothing_and_ignore_match_oracle():
    from audio_analysis_b200 import decay, rt60bands as rb
    g = load_golden("smooth_60000")
    ds = decay.DecayAnalysisSettings(edc_smoothing_window_samples=21, ignore_leading_seconds=0.005)
    s = rb.Rt60BandsAnalysisSettings(band_mode="octave", include_t20=True, decay_settings=ds)
    r = rb.analyse_rt60_bands_for_channel(g["x"], SR, "m", s)
    w = oracle.band_metrics(g["x"], SR, oracle.BandParams(
        band_mode="octave", include_t20=True,
        decay=oracle.DecayParams(edc_smoothing_window_samples=21, ignore_leading_seconds=0.005)))
    names = [b["name"] for b in w["bands"]]
    want = np.array([[np.nan if v is None else v for v in w["metrics"][n]] for n in names])
    have = np.array([[np.nan if v is None else v for v in
                      (r.band_metrics_by_name[n].rt60_t30_seconds, r.band_metrics_by_name[n].rt60_t20_seconds,
                       r.band_metrics_by_name[n].edt_seconds)] for n in names])
    check_band_table("smooth+ignore", have, want)


def test_rt60bands_errors():
    from audio_analysis_b200 import rt60bands as rb
    with pytest.raises(ValueError, match="Not enough samples"):
        rb.analyse_rt60_bands_for_channel(np.ones(7, np.float32), SR, "m", rb.Rt60BandsAnalysisSettings())
    with pytest.raises(ValueError, match="Unknown band_mode"):
        rb.analyse_rt60_bands_for_channel(np.ones(64, np.float32), SR, "m",
                                          rb.Rt60BandsAnalysisSettings(band_mode="quarter"))


# ------------------------------------------------------------------------------------ waterfall


def _check_slices(tag, got, want):
    assert got.shape == want.shape, (tag, got.shape, want.shape)
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    note("csd_rel_db_max_abs_err", err)
    note(f"csd_rel_db_max_abs_err[{tag}]", err)
    assert err <= CSD_TOL_DB, f"{tag}: CSD slices differ by {err} dB (> {CSD_TOL_DB})"


def test_waterfall_matches_reference_fixtures(golden):
    from audio_analysis_b200 import waterfall as wf
    x = golden["x"]
    r = wf.analyse_waterfall_for_channel(x, SR, "mono", wf.WaterfallAnalysisSettings())
    assert r.analysis_start_sample_index == int(golden["wf_start"])
    assert r.analysis_length_samples == int(golden["wf_length"])
    assert np.array_equal(r.slice_times_seconds, golden["wf_times"])
    assert np.array_equal(r.frequency_hz, golden["wf_freq"])
    _check_slices(golden["name"], r.slice_magnitude_rel_db, golden["wf_rel_db"])
    r2 = wf.analyse_waterfall_for_channel(x, SR, "mono", wf.WaterfallAnalysisSettings(
        db_reference="slice_max", num_slices=7, slice_mode="uniform_frames", dynamic_range_db=60.0))
    assert np.array_equal(r2.slice_times_seconds, golden["wf2_times"])
    _check_slices(golden["name"] + "/slice_max", r2.slice_magnitude_rel_db, golden["wf2_rel_db"])
    r3 = wf.analyse_waterfall_for_channel(x, SR, "mono", wf.WaterfallAnalysisSettings(
        smoothing_log_bins=9, slice_mode="uniform_time", slice_spacing_seconds=0.08))
    assert np.array_equal(r3.slice_times_seconds, golden["wf3_times"])
    _check_slices(golden["name"] + "/smoothed", r3.slice_magnitude_rel_db, golden["wf3_rel_db"])


def test_waterfall_other_fft_sizes_and_batch():
    from audio_analysis_b200 import waterfall as wf
    from audio_analysis_b200.synth import synth_batch_numpy
    x = synth_batch_numpy(5, 144000, SR)
    for n_fft, hop, hann in ((1024, 256, True), (8192, 1024, True), (2048, 512, False)):
        s = wf.WaterfallAnalysisSettings(n_fft=n_fft, hop_length=hop, use_hann_window=hann, num_slices=12)
        got = wf.analyse_waterfall_batch(x, SR, None, s)
        p = oracle.CsdParams(n_fft=n_fft, hop_length=hop, use_hann_window=hann, num_slices=12)
        for i in range(x.shape[0]):
            w = oracle.waterfall(x[i], SR, p)
            assert got[i].analysis_start_sample_index == w["start"]
            assert np.array_equal(got[i].slice_times_seconds, w["slice_times"])
            assert np.array_equal(got[i].frequency_hz, w["freq"])
            _check_slices(f"nfft{n_fft}", got[i].slice_magnitude_rel_db, w["rel_db"])


def test_waterfall_errors():
    from audio_analysis_b200 import waterfall as wf
    with pytest.raises(ValueError, match="need at least n_fft"):
        wf.analyse_waterfall_for_channel(np.ones(4000, np.float32), SR, "m", wf.WaterfallAnalysisSettings())
    x = load_golden("smooth_48000")["x"]
    with pytest.raises(ValueError, match="frequency selection is empty"):
        wf.analyse_waterfall_for_channel(x, SR, "m", wf.WaterfallAnalysisSettings(f_min_hz=5.0, f_max_hz=6.0))
    with pytest.raises(ValueError, match="Not enough slices"):
        wf.analyse_waterfall_for_channel(x, SR, "m", wf.WaterfallAnalysisSettings(analysis_duration_seconds=0.09))


def test_band_decimation_against_oracle_and_switch():
    """N = 288000 / 192000 rows: the decimated band path (default) and the full-rate path both sit inside the RT60
    tolerance of the oracle; aa_bands_set_decimation switches between them; calls that ask for T20 or EDT keep the
    decimation at <= 4 (csrc/bands.cu: fit_decimation_limit)."""
    from audio_analysis_b200 import rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_numpy
    lib = nat_lib()
    try:
        for n in (288000, 192000):
            x = synth_batch_numpy(2, n, SR, first_index=40)
            want = []
            for i in range(2):
                ob = oracle.band_metrics(x[i], SR, oracle.BandParams(band_mode="third", include_t20=True, include_edt=True))
                want.append(np.array([[np.nan if v is None else v for v in ob["metrics"][b["name"]]] for b in ob["bands"]]))
            names = [b["name"] for b in ob["bands"]]
            s30 = rb.Rt60BandsAnalysisSettings(band_mode="third")
            bands = rb._build_band_definitions(s30, SR)
            got = {}
            for max_d in (24, 16, 8, 4, 2, 1):
                lib.aa_bands_set_decimation(max_d)
                assert lib.aa_bands_get_decimation() == max_d
                dec = rb.decimation_classes(bands, n, SR, s30.transition_width_octaves, settings=s30)
                assert max(dec) <= max_d and (max_d == 1 or max(dec) > 1)
                res = rb.analyse_rt60_bands_batch(x, SR, ["a", "b"], s30)
                tab = np.array([_band_table(r, names) for r in res])
                got[max_d] = tab
                for i in range(2):
                    check_band_table(f"decim{max_d}/n{n}", tab[i][:, :1], want[i][:, :1])
                    err = np.nanmax(np.abs(tab[i][:, 0] - want[i][:, 0]) / want[i][:, 0])
                    note(f"band_t30_rel_err[decimation<={max_d}]", err)
            assert not np.array_equal(got[16], got[1])     # the switch really changes the path
            assert not np.array_equal(got[16], got[4]) and not np.array_equal(got[24], got[16])
            # T20 (20 dB) and EDT (10 dB from 0 dB) follow the staircase of a low band's curve more closely: such
            # calls stay at D <= 4 whatever the limit, and every metric stays inside the tolerance
            lib.aa_bands_set_decimation(16)
            for se in (rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True),
                       rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)):
                assert max(rb.decimation_classes(bands, n, SR, se.transition_width_octaves, settings=se)) == 4
                res = rb.analyse_rt60_bands_batch(x, SR, ["a", "b"], se)
                cols = 3 if se.include_edt else 2
                for i in range(2):
                    check_band_table(f"decim_t20_edt/n{n}", _band_table(res[i], names)[:, :cols], want[i][:, :cols])
            assert max(rb.decimation_classes(bands, n, SR, s30.transition_width_octaves, settings=s30)) == 16
    finally:
        lib.aa_bands_set_decimation(16)


def test_verb_sized_stereo_file_configs_1_and_2(dec):
    """BASELINE.json configs 1-2 at their real shape (plots/example/verb_report.md:6): a stereo file of 1,151,844
    samples per channel = 2^2 * 3 * 95987, so the band filter's circular FFT (rt60bands.py:170-175, over the WHOLE
    file, :345-350) goes through Bluestein at 2^21 and the EDC runs at 4x the bench's row length.
    config 1 = rt60bands --mono third-octave; config 2 = decay + waterfall per channel; fixture from the unmodified
    reference (tests/golden/make_golden_verb.py)."""
    from audio_analysis_b200 import rt60bands as rb, waterfall as wf
    g = load_verb_case()
    n = int(g["n"])
    assert n == 1151844 and g["x"].shape == (n, 2)
    for c, name in enumerate(("left", "right")):
        x = np.ascontiguousarray(g["x"][:, c])
        r = dec.analyse_decay_for_channel(x, SR, name, dec.DecayAnalysisSettings(compute_edt=True))
        assert r.analysis_start_sample_index == int(g[f"{name}_decay_start"])
        assert r.edc_db.size == int(g[f"{name}_decay_len"])
        check_edc(f"verb/{name}", r.edc_db[::64], g[f"{name}_decay_edc_db_dec64"])
        check_fit_rows(f"verb/{name}", _fits_table(r.fits), g[f"{name}_decay_fits"])
        assert r.early_decay_10db_time_seconds == pytest.approx(float(g[f"{name}_decay_early10"]), abs=1e-4)
        w = wf.analyse_waterfall_for_channel(x, SR, name, wf.WaterfallAnalysisSettings())
        assert w.analysis_start_sample_index == int(g[f"{name}_wf_start"])
        assert w.analysis_length_samples == int(g[f"{name}_wf_length"])
        assert np.array_equal(w.slice_times_seconds, g[f"{name}_wf_times"])
        assert np.array_equal(w.frequency_hz, g["wf_freq"])
        _check_slices(f"verb/{name}", w.slice_magnitude_rel_db, g[f"{name}_wf_rel_db"])
    s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    r = rb.analyse_rt60_bands_for_channel(g["mono"], SR, "mono", s)
    names = [b.name for b in r.band_definitions]
    assert names == [str(v) for v in g["mono_bands_third_names"]]
    check_band_table("verb/mono/third", _band_table(r, names), g["mono_bands_third_metrics"])
    # the stereo file as one batch of two rows (the file-level entry points loop channels; the batch path must agree)
    both = rb.analyse_rt60_bands_batch(np.ascontiguousarray(g["x"].T), SR, ["left", "right"],
                                       rb.Rt60BandsAnalysisSettings(band_mode="third"))
    for c in range(2):
        ob = oracle.band_metrics(np.ascontiguousarray(g["x"][:, c]), SR, oracle.BandParams(band_mode="third"))
        want = np.array([[np.nan if ob["metrics"][nm][0] is None else ob["metrics"][nm][0]] for nm in names])
        check_band_table(f"verb/ch{c}/third", _band_table(both[c], names)[:, :1], want)


def _band_tail(got: np.ndarray, ref: np.ndarray):
    rel = np.abs(got - ref) / np.abs(ref)
    outside = (rel > FIT_REL_TOL) & (np.abs(got - ref) > FIT_ABS_TOL_S)
    return rel, int(outside.sum()), int((np.isnan(got) != np.isnan(ref)).sum())


@pytest.mark.parametrize("n", [288000, 192000])
def test_band_certification_at_scale(n):
    """The DEFAULT band path (decimated synthesis, certified per unit, uncertified units redone at D <= 4 and then at
    full rate; csrc/bands.cu, bandfit.cu) against the full-rate device path - the reference's estimator,
    rt60bands.py:272-321 + decay.py:173-260 - over 4096 rows x 26 bands: NO metric outside 0.5 % / 1 ms, no None
    mismatch; and 8 of the rows against the oracle."""
    import torch
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    rows_n = 4096
    x = synth_batch_torch(rows_n, n, SR, device="cuda", first_index=9000)
    rows = engine.make_rows(x)
    s = rb.Rt60BandsAnalysisSettings(band_mode="third")
    ref = rb.run_rt60_bands(rows, SR, s, options=engine.bands_options(max_decimation=1))[0].cpu().numpy()[:, :, 0]
    assert engine.last_bands_stats["limits"] == [1]
    got = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()[:, :, 0]
    stats = dict(engine.last_bands_stats)
    assert stats["limits"] == [16, 8, 4, 1] and stats["units"][0] == rows_n * 26
    redo = sum(stats["units"][1:]) / float(rows_n * 26)
    note(f"band_certification_redo_fraction[n{n}]", redo)
    assert 0.0 < redo < 0.2, stats
    rel, outside, nan_mismatch = _band_tail(got, ref)
    note(f"band_t30_rel_err_certified_vs_full_rate[n{n}]", np.nanmax(rel))
    assert got.size >= 100000 and outside == 0 and nan_mismatch == 0, (outside, nan_mismatch, float(np.nanmax(rel)))
    assert np.nanmax(rel) <= 3.5e-3    # observed <= 1.5e-3 over 430 k metrics; the bar is 5e-3
    # the first level alone (round-1 behaviour) is a different, uncertified result
    raw = rb.run_rt60_bands(rows, SR, s, options=engine.bands_options(refine=False))[0].cpu().numpy()[:, :, 0]
    assert engine.last_bands_stats["limits"] == [16] and not np.array_equal(raw, got)
    # rows against the oracle (full float64 reference arithmetic)
    xs = x[:8].cpu().numpy()
    for i in range(8):
        ob = oracle.band_metrics(xs[i], SR, oracle.BandParams(band_mode="third"))
        want = np.array([[np.nan if ob["metrics"][b["name"]][0] is None else ob["metrics"][b["name"]][0]] for b in ob["bands"]])
        check_band_table(f"certified/n{n}/row{i}", got[i][:, None], want)
    del x, rows
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [288000, 192000])
def test_fused_band_kernel_against_the_three_kernel_path(n, monkeypatch):
    """band_fused_kernel (csrc/bandfused.cu: masked spectrum -> in-shared-memory inverse -> Schroeder sum -> fits in one
    CTA, for the classes with n / 2D <= 12,000) against the columns / rows / band_fit kernels it replaces
    (AA_NO_FUSED_BANDS): the same decimated estimator, so the uncertified first level agrees to rounding, and the
    certified default path stays inside the tolerance against the full-rate path (rt60bands.py:272-321)."""
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    rows_n = 256
    rows = engine.make_rows(synth_batch_torch(rows_n, n, SR, device="cuda", first_index=31000))
    s = rb.Rt60BandsAnalysisSettings(band_mode="third")
    raw_opt = engine.bands_options(refine=False)
    monkeypatch.delenv("AA_NO_FUSED_BANDS", raising=False)
    raw_f = rb.run_rt60_bands(rows, SR, s, options=raw_opt)[0].cpu().numpy()[:, :, 0]
    got_f = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()[:, :, 0]
    stats_f = dict(engine.last_bands_stats)
    monkeypatch.setenv("AA_NO_FUSED_BANDS", "1")
    raw_3 = rb.run_rt60_bands(rows, SR, s, options=raw_opt)[0].cpu().numpy()[:, :, 0]
    got_3 = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()[:, :, 0]
    stats_3 = dict(engine.last_bands_stats)
    monkeypatch.delenv("AA_NO_FUSED_BANDS", raising=False)
    assert np.array_equal(np.isnan(raw_f), np.isnan(raw_3))
    rel = np.abs(raw_f - raw_3) / np.abs(raw_3)
    note(f"band_fused_vs_three_kernel_first_level_rel[n{n}]", np.nanmax(rel))
    # same grid, same estimator; the transforms round differently (1e-6 of the signal) and the tiles are 128 samples
    # instead of 2 M1, so a crossing can move by one decimated sample on a flat stretch of the curve
    assert np.nanmedian(rel) <= 2e-6 and np.nanmax(rel) <= 2e-3, (float(np.nanmedian(rel)), float(np.nanmax(rel)))
    assert np.nanmean(rel > 1e-4) <= 2e-3
    # certified default: both paths inside the tolerance against the full-rate estimator, nearly the same units redone
    ref = rb.run_rt60_bands(rows, SR, s, options=engine.bands_options(max_decimation=1))[0].cpu().numpy()[:, :, 0]
    for got in (got_f, got_3):
        rel, outside, nan_mismatch = _band_tail(got, ref)
        assert outside == 0 and nan_mismatch == 0 and np.nanmax(rel) <= 3.5e-3
    assert stats_f["limits"] == stats_3["limits"] and stats_f["units"][0] == stats_3["units"][0]
    assert abs(sum(stats_f["units"][1:]) - sum(stats_3["units"][1:])) <= 0.02 * sum(stats_3["units"][1:]) + 8


def test_band_certification_with_t20_and_edt():
    """Calls that also ask for T20 / EDT start at D <= 4 and certify every fit (shorter fits need more points)."""
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    n, rows_n = 288000, 1024
    rows = engine.make_rows(synth_batch_torch(rows_n, n, SR, device="cuda", first_index=777))
    s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    ref = rb.run_rt60_bands(rows, SR, s, options=engine.bands_options(max_decimation=1))[0].cpu().numpy()
    got = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()
    assert engine.last_bands_stats["limits"] == [4, 1]
    for c, name in enumerate(("T30", "T20", "EDT")):
        rel, outside, nan_mismatch = _band_tail(got[:, :, c], ref[:, :, c])
        note(f"band_{name.lower()}_rel_err_certified_vs_full_rate", np.nanmax(rel))
        assert outside == 0 and nan_mismatch == 0, (name, outside, nan_mismatch)


def test_band_workspace_chunks_are_invisible():
    """aa_rt60bands_batch processes the rows in chunks that fit the caller's workspace (the bench runs 3 chunks per
    8192 rows): a workspace that forces >= 3 chunks gives bit-identical tables, decimated + certified or full rate."""
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    n, rows_n = 288000, 200
    rows = engine.make_rows(synth_batch_torch(rows_n, n, SR, device="cuda", first_index=31))
    s = rb.Rt60BandsAnalysisSettings(band_mode="third")
    plan = engine.bands_plan(n, rows.x.device)
    one_row = int(nat_lib().aa_bands_workspace_bytes(plan.handle, 1, 26))   # the smallest workspace the call accepts
    for opt in (None, engine.bands_options(max_decimation=1)):
        whole = rb.run_rt60_bands(rows, SR, s, options=opt)[0].cpu().numpy()
        units_whole = list(engine.last_bands_stats["units"])
        assert engine.last_bands_stats["chunks"] == 1
        small = rb.run_rt60_bands(rows, SR, s, options=opt, workspace_bytes=one_row)[0].cpu().numpy()
        assert engine.last_bands_stats["chunks"] >= 3, engine.last_bands_stats
        assert np.array_equal(whole, small, equal_nan=True)
        assert engine.last_bands_stats["units"] == units_whole


# ------------------------------------------------------------------------------------ spectrogram


def _spec_golden(case: str, variant: str) -> dict:
    from conftest import GOLDEN_DIR
    with np.load(GOLDEN_DIR / "spectrogram.npz") as z:
        return {k.split("/", 2)[2]: z[k] for k in z.files if k.startswith(f"{case}/{variant}/")}


@pytest.mark.parametrize("variant", ["default", "n1024", "n256"])
def test_spectrogram_matches_reference_fixtures(golden, variant):
    """Dense STFT (TMA-staged frame runs for n_fft >= 1024, the slice kernel below) against the reference's own
    output: stored columns within the CSD tolerance, every other frame through its column sum."""
    from audio_analysis_b200 import spectrogram as sp
    settings = {
        "default": sp.SpectrogramAnalysisSettings(),
        "n1024": sp.SpectrogramAnalysisSettings(n_fft=1024, hop_length=256, use_hann_window=False,
                                                ignore_leading_seconds=0.01, analysis_duration_seconds=0.5,
                                                floor_db=-100.0),
        "n256": sp.SpectrogramAnalysisSettings(n_fft=256, hop_length=200, trim_to_peak=False),
    }[variant]
    g = _spec_golden(golden["name"], variant)
    r = sp.analyse_spectrogram_for_channel(golden["x"], SR, "mono", settings)
    assert r.analysis_start_sample_index == int(g["start"]) and r.analysis_length_samples == int(g["length"])
    assert r.magnitude_db.shape == (settings.n_fft // 2 + 1, int(g["n_frames"]))
    assert r.magnitude_db.dtype == np.float32 and r.time_seconds.dtype == np.float32
    assert r.time_seconds[-1] == g["time_last"]
    assert np.array_equal(r.frequency_hz, np.fft.rfftfreq(settings.n_fft, 1.0 / SR).astype(np.float32))
    err = float(np.max(np.abs(r.magnitude_db[:, g["cols"]].astype(np.float64) - g["mag_db_cols"])))
    note(f"spectrogram_db_max_abs_err[{variant}]", err)
    # bins at the floor sit on a cliff of the log; compare those in linear magnitude instead
    lin_g = 10.0 ** (r.magnitude_db[:, g["cols"]].astype(np.float64) / 20.0)
    lin_w = 10.0 ** (g["mag_db_cols"].astype(np.float64) / 20.0)
    scale = lin_w.max(axis=0, keepdims=True)
    rel = float(np.max(np.abs(lin_g - lin_w) / np.maximum(lin_w, 1e-4 * scale)))
    note("spectrogram_lin_rel_err", rel)
    assert rel <= 1e-4 or err <= CSD_TOL_DB, (variant, rel, err)
    sums = r.magnitude_db.astype(np.float64).sum(axis=0)
    assert np.max(np.abs(sums - g["col_sums"])) <= 0.05 * r.magnitude_db.shape[0] ** 0.5
    assert sp.summarise_spectrogram_results_text([r]) == str(g["summary"])


def test_spectrogram_batch_oracle_and_errors():
    from audio_analysis_b200 import spectrogram as sp
    rng = np.random.default_rng(5)
    n = 70000
    chans = np.stack([rng.standard_normal(n).astype(np.float32) * np.exp(-np.arange(n) / (4000.0 * (i + 1))).astype(np.float32)
                      for i in range(3)])
    s = sp.SpectrogramAnalysisSettings(n_fft=2048, hop_length=300, ignore_leading_seconds=0.02)
    res = sp.analyse_spectrogram_batch(chans, SR, ["a", "b", "c"], s)
    for i, r in enumerate(res):
        o = oracle.spectrogram(chans[i], SR, oracle.SpectrogramParams(n_fft=2048, hop_length=300,
                                                                      ignore_leading_seconds=0.02))
        assert r.analysis_start_sample_index == o["start"] and r.magnitude_db.shape == o["mag_db"].shape
        lin_g = 10.0 ** (r.magnitude_db.astype(np.float64) / 20.0)
        lin_w = 10.0 ** (o["mag_db"].astype(np.float64) / 20.0)
        rel = float(np.max(np.abs(lin_g - lin_w) / np.maximum(lin_w, 1e-4 * lin_w.max())))
        note("spectrogram_lin_rel_err", rel)
        assert rel <= 1e-4
    with pytest.raises(ValueError, match="need at least n_fft"):
        sp.analyse_spectrogram_for_channel(np.ones(3000, np.float32), SR, "m", sp.SpectrogramAnalysisSettings())
    with pytest.raises(ValueError, match="expects a 1D mono array"):
        sp.analyse_spectrogram_for_channel(np.ones((2, 5000), np.float32), SR, "m", sp.SpectrogramAnalysisSettings())


# ------------------------------------------------------------------------------------ modal cloud


def _modal_compare(tag, got_points, want: np.ndarray):
    """Points are matched by bin centre.  A bin's curve is a noisy STFT magnitude, so "first frame <= target" can
    move by a frame when a value differs in the 4th decimal of a dB: most points agree to 1e-5, a few move."""
    got = {round(p.centre_hz, 3): p for p in got_points}
    ref = {round(float(c), 3): (float(rt), float(r2)) for c, rt, r2 in want}
    both = sorted(set(got) & set(ref))
    only = len(set(got) ^ set(ref))
    rel = np.array([abs(got[c].rt60_seconds - ref[c][0]) / ref[c][0] for c in both])
    note(f"modal_points_missing_or_extra_frac[{tag}]", only / max(1, len(ref)))
    note(f"modal_rt60_rel_err_median[{tag}]", float(np.median(rel)) if rel.size else 0.0)
    note(f"modal_rt60_frac_outside_0.5pct[{tag}]", float(np.mean(rel > FIT_REL)) if rel.size else 0.0)
    assert only <= max(2, 0.03 * len(ref)), (tag, only, len(ref))
    assert rel.size and float(np.median(rel)) <= 1e-4, (tag, float(np.median(rel)))
    assert float(np.mean(rel > FIT_REL)) <= 0.03, (tag, float(np.mean(rel > FIT_REL)))
    r2 = np.array([abs(got[c].r_squared - ref[c][1]) for c in both])
    assert float(np.median(r2)) <= 1e-4


FIT_REL = 0.005


@pytest.mark.parametrize("case,variant", [("smooth_48000", "n2048"), ("odd_45001", "n2048_t20"),
                                          ("prime_factor_50022", "n1024_edt"), ("long_96000", "default"),
                                          ("long_144000", "default"), ("long_144000", "n2048_t20")])
def test_modal_cloud_matches_reference_fixtures(case, variant):
    from conftest import GOLDEN_DIR
    from audio_analysis_b200 import modalcloud as mc
    with np.load(GOLDEN_DIR / "modalcloud.npz") as z:
        g = {k: z[k] for k in z.files}
    pcm = g[f"{case}/pcm"] if f"{case}/pcm" in g else load_golden(case)["pcm"]
    x = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
    settings = {
        "default": mc.ModalCloudAnalysisSettings(),
        "n2048": mc.ModalCloudAnalysisSettings(n_fft=2048, hop_length=128),
        "n2048_t20": mc.ModalCloudAnalysisSettings(n_fft=2048, hop_length=128, metric="t20", log_bins_per_octave=12,
                                                   ignore_leading_seconds=0.005, min_fit_points=6),
        "n1024_edt": mc.ModalCloudAnalysisSettings(n_fft=1024, hop_length=64, metric="edt", f_min_hz=100.0,
                                                   f_max_hz=12000.0, use_hann_window=False),
    }[variant]
    key = f"{case}/{variant}"
    r = mc.analyse_modal_cloud_for_channel(x, SR, "mono", settings)
    assert r.analysis_start_sample_index == int(g[f"{key}/start"]) and r.analysis_length_samples == int(g[f"{key}/length"])
    assert r.metric == str(g[f"{key}/metric"])
    _modal_compare(key, r.points, g[f"{key}/points"])
    text = mc.summarise_modal_cloud_results_text([r])
    assert text.splitlines()[0].rsplit(" points=", 1)[0] == str(g[f"{key}/summary"]).splitlines()[0].rsplit(" points=", 1)[0]


def test_modal_cloud_batch_and_errors():
    from audio_analysis_b200 import modalcloud as mc
    rng = np.random.default_rng(11)
    n = 120000
    chans = np.stack([(rng.standard_normal(n) * 10.0 ** (-3.0 * np.arange(n) / (SR * (0.6 + 0.3 * i)))).astype(np.float32)
                      for i in range(2)])
    s = mc.ModalCloudAnalysisSettings(n_fft=4096, hop_length=256)
    res = mc.analyse_modal_cloud_batch(chans, SR, ["a", "b"], s)
    for i, r in enumerate(res):
        o = oracle.modal_cloud(chans[i], SR, oracle.ModalParams(n_fft=4096, hop_length=256))
        assert r.analysis_start_sample_index == o["start"] and r.metric == "t30"
        _modal_compare(f"batch{i}", r.points, np.array(o["points"]).reshape(-1, 3))
        rt = np.median([p.rt60_seconds for p in r.points])
        assert abs(rt - (0.6 + 0.3 * i)) < 0.15 * (0.6 + 0.3 * i)     # the clouds sit at the synthetic RT60
    with pytest.raises(ValueError, match="need at least n_fft"):
        mc.analyse_modal_cloud_for_channel(np.ones(5000, np.float32), SR, "m", mc.ModalCloudAnalysisSettings())


# ------------------------------------------------------------------------------------ BASELINE.json configs 3 and 4


def _band_table(result, names):
    return np.array([[np.nan if v is None else v for v in
                      (result.band_metrics_by_name[n].rt60_t30_seconds, result.band_metrics_by_name[n].rt60_t20_seconds,
                       result.band_metrics_by_name[n].edt_seconds)] for n in names])


def test_config3_bundle_taps_pcm16_stereo():
    """Config 3 shape: stereo 4 s @48 kHz taps as interleaved PCM16 (recorder.hpp layout), decoded on the device
    (io.py:46-64 rules), then decay + third-octave rt60bands per channel - against the oracle on the same PCM."""
    import torch
    from audio_analysis_b200 import decay, engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_numpy
    taps, n = 3, 192000
    x = synth_batch_numpy(2 * taps, n, SR, first_index=300)
    pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16).reshape(taps, 2, n).transpose(0, 2, 1)
    rows = engine.pcm16_to_rows(torch.from_numpy(np.ascontiguousarray(pcm).reshape(taps, -1)).cuda(), n, 2,
                                mono_downmix=False)
    ds = decay.DecayAnalysisSettings(compute_edt=True)
    bs = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True, decay_settings=ds)
    dec_out = decay.run_decay(rows, SR, ds)
    got_bands = rb.analyse_rt60_bands_batch(rows, SR, None, bs)
    fits = dec_out.fits.cpu().numpy()
    starts = dec_out.start.cpu().numpy()
    for r in range(2 * taps):
        xr = (pcm[r // 2, :, r % 2].astype(np.float32) / 32768.0).astype(np.float32)
        w = oracle.decay_metrics(xr, SR, oracle.DecayParams(compute_edt=True))
        assert int(starts[r]) == w["start"]
        for c, name in enumerate(("EDT", "T20", "T30")):
            if w["fits"].get(name) is not None:
                assert seconds_close(fits[r, c, 5], w["fits"][name]["rt60"]), (r, name)
        wb = oracle.band_metrics(xr, SR, oracle.BandParams(band_mode="third", include_t20=True, include_edt=True))
        names = [b["name"] for b in wb["bands"]]
        want = np.array([[np.nan if v is None else v for v in wb["metrics"][nm]] for nm in names])
        check_band_table(f"config3/row{r}", _band_table(got_bands[r], names), want)


def test_config4_sweep_capture_to_bands_and_csd():
    """Config 4 chain at its real size: a 13 s capture of a 10 s log sweep through a synthetic room is deconvolved at
    n_fft = 2^20 (deconvolve.py:124-193), and the 624,000-sample IR goes through third-octave rt60bands and the CSD
    waterfall - every stage against the oracle chain on the same capture."""
    import torch
    from audio_analysis_b200 import engine, rt60bands as rb, waterfall as wf
    from audio_analysis_b200.synth import synth_batch_numpy
    n, n_sweep, caps = 624000, 480000, 2
    t = np.arange(n_sweep) / SR
    f0, f1, dur = 20.0, 20000.0, n_sweep / SR
    k = np.log(f1 / f0)
    sweep = np.zeros(n, np.float32)
    sweep[:n_sweep] = (0.5 * np.sin(2.0 * np.pi * f0 * dur / k * (np.exp(t / dur * k) - 1.0))).astype(np.float32)
    irs = synth_batch_numpy(caps, 96000, SR, first_index=900)
    S = np.fft.rfft(sweep.astype(np.float64))
    rec = np.stack([np.fft.irfft(S * np.fft.rfft(irs[i].astype(np.float64), n=n), n=n) for i in range(caps)])
    rec = (0.9 * rec / np.max(np.abs(rec))).astype(np.float32)
    n_fft = 1 << (n - 1).bit_length()
    assert n_fft == 1 << 20
    out = engine.deconvolve(engine.make_rows(rec), torch.from_numpy(sweep), n_fft, 1e-10, True, 0.95, 1, n)
    got_ir = out[:, :n].contiguous()
    bs = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True)
    got_bands = rb.analyse_rt60_bands_batch(engine.make_rows(got_ir), SR, None, bs)
    got_wf = wf.analyse_waterfall_batch(engine.make_rows(got_ir), SR, None, wf.WaterfallAnalysisSettings())
    ir_host = got_ir.cpu().numpy()
    for i in range(caps):
        want_ir = oracle.deconvolve(rec[i].reshape(-1, 1), sweep)[:, 0]
        err = float(np.max(np.abs(ir_host[i] - want_ir)))
        note("config4_deconvolve_max_abs_err", err)
        assert err <= 2e-4 * float(np.max(np.abs(want_ir)))
        wb = oracle.band_metrics(want_ir, SR, oracle.BandParams(band_mode="third", include_t20=True))
        names = [b["name"] for b in wb["bands"]]
        want = np.array([[np.nan if v is None else v for v in wb["metrics"][nm]] for nm in names])
        check_band_table(f"config4/capture{i}", _band_table(got_bands[i], names)[:, :2], want[:, :2])
        # CSD: stage parity on the SAME impulse response (the device one).  Chained through the two deconvolutions the
        # slices differ by ~2e-3 dB in the -70 dB bins, which is the float32-vs-float64 deconvolution error above,
        # not the STFT's.
        ww = oracle.waterfall(ir_host[i], SR, oracle.CsdParams())
        assert oracle.waterfall(want_ir, SR, oracle.CsdParams())["start"] == ww["start"]
        assert got_wf[i].analysis_start_sample_index == ww["start"]
        assert np.array_equal(got_wf[i].slice_times_seconds, ww["slice_times"])
        _check_slices(f"config4/capture{i}", got_wf[i].slice_magnitude_rel_db, ww["rel_db"])


# ------------------------------------------------------------------------------------ size-independent properties


def test_full_size_properties_config5_rows():
    """Config-5 row shape (288,000 samples): properties that do not need the oracle at full batch size."""
    import torch
    from audio_analysis_b200 import decay, engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    x = synth_batch_torch(64, 288000, SR, device="cuda")
    rows = engine.make_rows(x)
    s = decay.DecayAnalysisSettings(compute_edt=True)
    out = decay.run_decay(rows, SR, s)
    torch.cuda.synchronize()
    start = out.start.cpu().numpy()
    edc = out.edc_db.cpu().numpy()
    # 1. peak index = argmax |x|; 2. curve starts at exactly 0 dB; 3. monotone non-increasing; 4. floor respected
    assert np.array_equal(start, x.abs().argmax(dim=1).cpu().numpy())
    for b in range(8):
        c = edc[b, start[b]:]
        assert c[0] == 0.0 and np.all(np.diff(c) <= 1e-6) and c.min() >= -120.0
    # 5. scaling invariance: EDC in dB and all fits do not depend on the input gain
    out2 = decay.run_decay(engine.make_rows(x * 0.25), SR, s)
    assert np.array_equal(out2.start.cpu().numpy(), start)
    a, b2 = out.fits.cpu().numpy()[:, :, 5], out2.fits.cpu().numpy()[:, :, 5]
    assert np.allclose(a, b2, rtol=1e-6, equal_nan=True)
    # 6. band filters partition the spectrum: third-octave band signals of one row sum (with the complementary
    #    low/high remainder) to the input only inside the covered range -> energy of the sum <= energy of x
    metrics, bands, plan, st, sig = rb.run_rt60_bands(engine.make_rows(x[:2]), SR,
                                                      rb.Rt60BandsAnalysisSettings(band_mode="third"),
                                                      want_band_signals=True)
    total = sig.sum(dim=1)
    assert float((total ** 2).sum()) <= 1.6 * float((x[:2] ** 2).sum())
    # 7. shard invariance (what the multi-GPU path relies on): analysing a sub-batch gives the same table
    whole = rb.run_rt60_bands(engine.make_rows(x[:3]), SR, rb.Rt60BandsAnalysisSettings(band_mode="third"))[0]
    sub = rb.run_rt60_bands(engine.make_rows(x[1:2]), SR, rb.Rt60BandsAnalysisSettings(band_mode="third"))[0]
    assert torch.equal(torch.nan_to_num(sub[0]), torch.nan_to_num(whole[1]))
    # 8. band decimation (default) against the full-rate synthesis (what `want_band_signals` forces): the decimated
    #    grid holds exact samples of the same band signals, only the Schroeder sum is taken on fewer points
    assert nat_lib().aa_bands_get_decimation() == 16
    full = metrics.cpu().numpy()[:2, :, 0]
    dec = whole.cpu().numpy()[:2, :, 0]
    assert np.array_equal(np.isnan(full), np.isnan(dec))
    rel = np.nanmax(np.abs(dec - full) / full)
    note("band_rt60_decimated_vs_full_rel", rel)
    assert rel <= 5e-3


def test_hot_path_batch_of_2048_config5_rows_sampled_against_the_oracle():
    """pipeline.run_hot_path at the bench's shape (rows of 288,000 samples, 2048 of them in one batch: one peak search,
    decay + third-octave bands + CSD): rows sampled across the batch are checked against the oracle, and the whole
    batch against the properties that need no oracle."""
    import torch
    from audio_analysis_b200 import decay, engine, rt60bands as rb, waterfall as wf
    from audio_analysis_b200.pipeline import run_hot_path
    from audio_analysis_b200.synth import synth_batch_torch
    rows_n, n = 2048, 288000
    x = synth_batch_torch(rows_n, n, SR, device="cuda", first_index=60000)
    ds = decay.DecayAnalysisSettings(compute_edt=True)
    bs = rb.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
    out = run_hot_path(engine.make_rows(x), SR, ds, bs, wf.WaterfallAnalysisSettings(), want_edc=True)
    torch.cuda.synchronize()
    start = out.decay.start.cpu().numpy()
    assert np.array_equal(start, x.abs().argmax(dim=1).cpu().numpy())
    assert int(out.decay.status.abs().sum()) == 0
    fits = out.decay.fits.cpu().numpy()
    bands = out.band_metrics.cpu().numpy()
    assert np.isfinite(fits[:, :, 5]).all() and np.isfinite(bands[:, :, 0]).all()
    csd = out.csd.cpu().numpy()
    assert csd.shape == (rows_n, 18, 1705) and csd.max() == 0.0 and csd.min() >= -80.0
    names = [b.name for b in out.bands]
    for i in (0, 1, 511, 1024, 1777, rows_n - 1):
        xi = x[i].cpu().numpy()
        od = oracle.decay_metrics(xi, SR, oracle.DecayParams(compute_edt=True))
        assert start[i] == od["start"]
        check_edc(f"batch2048/row{i}", out.decay.edc_db[i, start[i]:n].cpu().numpy(), od["edc_db"])
        for r, name in enumerate(("EDT", "T20", "T30")):
            want = od["fits"][name]["rt60"]
            assert seconds_close(fits[i, r, 5], want), (i, name, fits[i, r, 5], want)
        ob = oracle.band_metrics(xi, SR, oracle.BandParams(band_mode="third"))
        want = np.array([[np.nan if ob["metrics"][nm][0] is None else ob["metrics"][nm][0]] for nm in names])
        check_band_table(f"batch2048/row{i}", bands[i][:, :1], want)
        ow = oracle.waterfall(xi, SR, oracle.CsdParams())
        _check_slices(f"batch2048/row{i}", csd[i][:ow["rel_db"].shape[0]], ow["rel_db"])
    del x
    torch.cuda.empty_cache()


def test_host_streaming_pipeline_equals_resident_batch():
    """run_hot_path_host (chunked H2D / kernels / D2H on three streams) returns exactly what the resident batch
    returns, for a chunk size that does not divide the batch."""
    import torch
    from audio_analysis_b200 import decay, engine, rt60bands as rb, waterfall as wf
    from audio_analysis_b200.pipeline import run_hot_path, run_hot_path_host
    from audio_analysis_b200.synth import synth_batch_torch
    n = 96000
    x = synth_batch_torch(11, n, SR, device="cuda")
    ds = decay.DecayAnalysisSettings(compute_edt=True)
    bs = rb.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
    ws = wf.WaterfallAnalysisSettings()
    ref = run_hot_path(engine.make_rows(x), SR, ds, bs, ws, want_edc=True)
    xh = torch.empty((11, n), dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    torch.cuda.synchronize()
    got = run_hot_path_host(xh, SR, ds, bs, ws, chunk_rows=4, want_edc=True)
    got.wait()
    torch.cuda.synchronize()
    assert got.h2d_bytes == 11 * n * 4 and got.d2h_bytes > 0
    assert torch.equal(got.start, ref.decay.start.cpu())
    assert torch.equal(torch.nan_to_num(got.fits), torch.nan_to_num(ref.decay.fits.cpu()))
    assert torch.equal(torch.nan_to_num(got.early10), torch.nan_to_num(ref.decay.early10.cpu()))
    assert torch.equal(torch.nan_to_num(got.band_rt60), torch.nan_to_num(ref.band_metrics.cpu()))
    assert torch.equal(got.csd, ref.csd.cpu())
    st = ref.decay.start.cpu().numpy()
    for b in range(11):
        assert torch.equal(got.edc_db[b, st[b]:n], ref.decay.edc_db[b, st[b]:n])
    # numpy input (pageable) goes through the same path
    got2 = run_hot_path_host(xh.numpy().copy(), SR, ds, bs, ws, chunk_rows=16, want_edc=False, do_csd=False)
    got2.wait()
    torch.cuda.synchronize()
    assert got2.edc_db is None and got2.csd is None
    assert torch.equal(torch.nan_to_num(got2.band_rt60), torch.nan_to_num(ref.band_metrics.cpu()))
    # 16-bit PCM host input: decoded on the device exactly like io.py:46-64 (sample / 32768)
    pcm = (x * 32767.0).round().clamp(-32768, 32767).to(torch.int16).cpu()
    want = run_hot_path(engine.make_rows(pcm.to(torch.float32).cuda() / 32768.0), SR, ds, bs, ws, want_edc=False)
    got3 = run_hot_path_host(pcm, SR, ds, bs, ws, chunk_rows=5, want_edc=False)
    got3.wait()
    torch.cuda.synchronize()
    assert got3.h2d_bytes == 11 * n * 2
    assert torch.equal(got3.start, want.decay.start.cpu())
    assert torch.equal(torch.nan_to_num(got3.band_rt60), torch.nan_to_num(want.band_metrics.cpu()))
    assert torch.equal(got3.csd, want.csd.cpu())


def test_stage_profiler_reports_kernel_times():
    """aa_profile_*: CUDA-event time per library stage (what bench.py's roofline objects are built from)."""
    import torch
    from audio_analysis_b200 import _native as nat, decay, engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    lib = nat.load()
    x = synth_batch_torch(8, 288000, SR, device="cuda")
    rows = engine.make_rows(x)
    lib.aa_profile_enable(1)
    out = decay.run_decay(rows, SR, decay.DecayAnalysisSettings())
    rb.run_rt60_bands(rows, SR, rb.Rt60BandsAnalysisSettings(band_mode="third"), start_idx=out.start)
    torch.cuda.synchronize()
    assert lib.aa_profile_collect() >= 4
    for name in (b"edc", b"fft_cols_band", b"fft_rows_band", b"band_fit"):
        assert lib.aa_profile_ms(name) > 0.0 and lib.aa_profile_scopes(name) >= 1, name
    assert lib.aa_profile_ms(b"no_such_stage") == 0.0
    lib.aa_profile_enable(0)
    decay.run_decay(rows, SR, decay.DecayAnalysisSettings())
    torch.cuda.synchronize()
    assert lib.aa_profile_collect() == 0


# ------------------------------------------------------------------------------------ PCM16, deconvolution, CLI


def test_pcm16_decode_matches_io_rules():
    import torch
    from audio_analysis_b200 import engine
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, size=(3, 1000, 2), dtype=np.int16)
    dev_pcm = torch.from_numpy(pcm.reshape(3, -1)).cuda()
    rows = engine.pcm16_to_rows(dev_pcm, 1000, 2, mono_downmix=False)
    want = (pcm.astype(np.float32) / 32768.0)
    got = rows.x.cpu().numpy()
    assert np.array_equal(got[0::2, :1000], want[:, :, 0]) and np.array_equal(got[1::2, :1000], want[:, :, 1])
    mono = engine.pcm16_to_rows(dev_pcm, 1000, 2, mono_downmix=True).x.cpu().numpy()
    assert np.array_equal(mono[:, :1000], 0.5 * (want[:, :, 0] + want[:, :, 1]))


def test_deconvolve_matches_reference_fixture():
    import torch
    from audio_analysis_b200 import engine
    g = load_golden("deconvolve")
    rec = g["recorded"]  # (N, C)
    rows = engine.make_rows(np.ascontiguousarray(rec.T))
    n = max(rec.shape[0], g["sweep"].size)
    n_fft = 1 << (n - 1).bit_length()
    out = engine.deconvolve(rows, torch.from_numpy(g["sweep"]), n_fft, 1e-10, True, 0.95, rec.shape[1], rec.shape[0])
    got = out.cpu().numpy()[:, : rec.shape[0]].T
    err = float(np.max(np.abs(got - g["ir"])))
    note("deconvolve_max_abs_err", err)
    assert err <= 2e-4 * float(np.max(np.abs(g["ir"])))


def test_cli_end_to_end(tmp_path, capsys):
    """`python -m analyse.cli ...` surface on a PCM16 stereo WAV: same text as the oracle-derived summary."""
    from scipy.io import wavfile
    from audio_analysis_b200 import cli
    a, b = load_golden("smooth_48000"), load_golden("onset_at_zero_48000")
    wav = tmp_path / "ir.wav"
    wavfile.write(str(wav), SR, np.stack([a["pcm"], b["pcm"]], axis=1))
    cli.main(["decay", "--input", str(wav), "--no_show"])
    text = capsys.readouterr().out
    assert f"[left] analysis_start_sample_index={int(a['decay_start'])}" in text
    assert f"[right] analysis_start_sample_index={int(b['decay_start'])}" in text
    assert "EDT: range=[0.0,-10.0]dB" in text and "T30: range=[-5.0,-35.0]dB" in text
    want = oracle.decay_metrics(a["x"], SR, oracle.DecayParams(compute_edt=True))["fits"]["T30"]["rt60"]
    line = [ln for ln in text.splitlines() if ln.strip().startswith("T30")][0]
    assert seconds_close(float(line.split("rt60=")[1].rstrip("s")), want)
    cli.main(["rt60bands", "--input", str(wav), "--mono", "--band_mode", "third", "--no_show"])
    text = capsys.readouterr().out
    assert text.startswith("[mono]\nBand  T30_RT60(s)\n39Hz  ") and "12699Hz  " in text
    cli.main(["waterfall", "--input", str(wav), "--no_show"])
    text = capsys.readouterr().out
    assert "slices=18  f_bins=1705" in text and text.count("\n") == 2
    cli.main(["modalcloud", "--input", str(wav), "--no_show", "--mono", "--n_fft", "2048", "--hop_length", "128"])
    text = capsys.readouterr().out
    assert text.startswith("[mono] metric=t30 start_sample=") and "  rt60: median=" in text
    cli.main(["spectrogram", "--input", str(wav), "--no_show", "--mono", "--n_fft", "2048"])
    text = capsys.readouterr().out
    mono = (0.5 * (a["x"] + b["x"])).astype(np.float32)
    want = oracle.spectrogram(mono, SR, oracle.SpectrogramParams(n_fft=2048))
    assert text.strip() == (f"[mono] start_sample={want['start']}  len_samples={want['length']}  "
                            f"dur={want['length'] / SR:.3f}s  stft(n_fft=2048, frames={want['mag_db'].shape[1]})")


def test_bundle_end_to_end(tmp_path):
    """recorder.hpp bundle layout -> per-tap reports + index, all taps in shared launches."""
    import json
    from scipy.io import wavfile
    from audio_analysis_b200.bundle import BundleRunSettings, run_bundle_report
    from audio_analysis_b200.report import ReportSettings
    from audio_analysis_b200.rt60bands import Rt60BandsAnalysisSettings
    root = tmp_path / "run"
    (root / "taps").mkdir(parents=True)
    a, b = load_golden("smooth_48000"), load_golden("onset_at_zero_48000")
    for tap in ("tap_000", "tap_001"):
        wavfile.write(str(root / "taps" / f"{tap}.wav"), SR, np.stack([a["pcm"], b["pcm"]], axis=1))
    (root / "meta.json").write_text(json.dumps({"sample_rate_hz": SR, "length_samples": 48000,
                                                "taps": ["tap_000", "tap_001"]}))
    rs = ReportSettings(rt60_bands_settings=Rt60BandsAnalysisSettings(band_mode="third"))
    index = run_bundle_report(root, BundleRunSettings(report_settings=rs))
    assert index == root / "reports" / "bundle_report.md"
    assert "- [tap_001](reports/tap_001/tap_001_report.md)" in index.read_text()
    md = (root / "reports" / "tap_000" / "tap_000_report.md").read_text()
    assert "## Decay / EDC" in md and "## RT60 by band" in md and "## Waterfall" in md and "12699Hz" in md
    assert f"[left] analysis_start_sample_index={int(a['decay_start'])}" in md
    assert "Blocks not run" in md and "Spectrogram" not in md.split("Blocks not run")[1]
    # section order of the reference's report (report.py:252-386)
    order = [md.index(h) for h in ("## Decay / EDC", "## RT60 by band", "## Spectrogram", "## Waterfall", "## Modal cloud")]
    assert order == sorted(order)
    assert "stft(n_fft=4096, frames=" in md and "] metric=t30 start_sample=" in md


def test_bundle_wire_path_equals_per_file_analysis_and_isolates_a_bad_tap(tmp_path):
    """run_bundle_report on recorder-format taps (PCM16, staged once, decoded on the device, one peak search per
    channel) writes, for every tap, exactly the text the per-file functions produce; a mixed batch (one float32 WAV)
    takes the loader path with the same result; a tap that is too short for the waterfall raises the reference's
    ValueError after the taps before it have been written (bundle.py:56-67)."""
    import json
    from scipy.io import wavfile
    from audio_analysis_b200 import decay, rt60bands as rb, waterfall as wf, wire
    from audio_analysis_b200.bundle import BundleRunSettings, run_bundle_report
    from audio_analysis_b200.report import ReportSettings, run_reports_batch
    from audio_analysis_b200.synth import synth_ir_numpy
    n = 48000
    taps = {}
    for i in range(3):
        ch = np.stack([synth_ir_numpy(900 + 2 * i + c, n, SR) for c in range(2)], axis=1)
        taps[f"tap_{i:03d}"] = np.round(ch * 32767.0).astype(np.int16)
    root = tmp_path / "run"
    (root / "taps").mkdir(parents=True)
    for name, pcm in taps.items():
        wire.write_pcm16_wav(root / "taps" / f"{name}.wav", pcm, SR)
    (root / "meta.json").write_text(json.dumps({"sample_rate_hz": SR, "length_samples": n, "taps": list(taps)}))
    rs = ReportSettings(rt60_bands_settings=rb.Rt60BandsAnalysisSettings(band_mode="third"), run_spectrogram=False,
                        run_modal_cloud=False)
    run_bundle_report(root, BundleRunSettings(report_settings=rs))
    for name, pcm in taps.items():
        md = (root / "reports" / name / f"{name}_report.md").read_text()
        path = root / "taps" / f"{name}.wav"
        assert decay.summarise_decay_results_text(decay.analyse_decay_from_wav_file(path)).strip() in md
        bands = rb.analyse_rt60_bands_from_wav_file(path, rb.Rt60BandsAnalysisSettings(band_mode="third"))
        assert rb.summarise_rt60_bands_results_text(bands, False, False).strip() in md
        assert wf.summarise_waterfall_results_text(wf.analyse_waterfall_from_wav_file(path)).strip() in md
        assert f"**Channels:** 2" in md and f"**Samples:** {n}" in md
    # the loader path (one file is float32) renders the same report for the PCM16 taps
    other = tmp_path / "float.wav"
    wavfile.write(str(other), SR, (taps["tap_000"].astype(np.float32) / 32768.0))
    outs = run_reports_batch([root / "taps" / "tap_001.wav", other], [tmp_path / "o" / "tap_001", tmp_path / "o" / "b"], rs)
    wired = (root / "reports" / "tap_001" / "tap_001_report.md").read_text()
    assert outs[0].summary_markdown.split("---", 1)[1] == wired.split("---", 1)[1]
    # error isolation: the second of three files is too short for n_fft 4096
    short = tmp_path / "short.wav"
    wire.write_pcm16_wav(short, taps["tap_000"][:3000], SR)
    with pytest.raises(ValueError, match="need at least n_fft"):
        run_reports_batch([root / "taps" / "tap_000.wav", short, root / "taps" / "tap_002.wav"],
                          [tmp_path / "e" / "a", tmp_path / "e" / "b", tmp_path / "e" / "c"], rs)
    assert (tmp_path / "e" / "a_report.md").exists() and not (tmp_path / "e" / "c_report.md").exists()


def test_deconvolve_api_and_cli(tmp_path):
    """analyse.deconvolve mirror (deconvolve.py:124-268) and the `deconvolve` subcommand (cli.py:1219-1246)."""
    from scipy.io import wavfile
    from audio_analysis_b200 import cli, deconvolve as dv
    g = load_golden("deconvolve")
    want = g["ir"]
    got = dv.deconvolve_impulse_response(g["recorded"], g["sweep"], SR, dv.DeconvolveSettings())
    assert got.shape == want.shape and got.dtype == np.float32
    err = float(np.max(np.abs(got - want)))
    note("deconvolve_api_max_abs_err", err)
    assert err <= 2e-4 * float(np.max(np.abs(want)))
    # settings variants against the oracle
    for st in (dv.DeconvolveSettings(normalise_peak=False), dv.DeconvolveSettings(remove_dc=False, target_peak=0.5),
               dv.DeconvolveSettings(output_length_mode="full_fft")):
        got = dv.deconvolve_impulse_response(g["recorded"], g["sweep"], SR, st)
        ref = oracle.deconvolve(g["recorded"], g["sweep"], st.regularization_relative, st.normalise_peak,
                                st.target_peak, st.remove_dc, st.output_length_mode == "full_fft")
        assert got.shape == ref.shape and float(np.max(np.abs(got - ref))) <= 2e-4 * float(np.abs(ref).max())
    with pytest.raises(ValueError, match="at least a few samples"):
        dv.deconvolve_impulse_response(np.zeros((4, 1), np.float32), g["sweep"], SR, dv.DeconvolveSettings())
    with pytest.raises(ValueError, match="Unknown output_length_mode"):
        dv.deconvolve_impulse_response(g["recorded"], g["sweep"], SR, dv.DeconvolveSettings(output_length_mode="x"))
    # files + CLI: same numbers, float32 WAV written next to the recording by default
    rec_path, sw_path = tmp_path / "capture.wav", tmp_path / "sweep.wav"
    pad = np.zeros((SR // 4, 2), np.float32)
    wavfile.write(str(rec_path), SR, np.concatenate([g["recorded"], pad]))
    wavfile.write(str(sw_path), SR, g["sweep"])
    import io as _io
    from contextlib import redirect_stdout
    buf = _io.StringIO()
    with redirect_stdout(buf):
        cli.main(["deconvolve", "--recorded_wav_file_path", str(rec_path), "--sweep_wav_file_path", str(sw_path)])
    out_path = tmp_path / "capture_ir.wav"
    assert f"Wrote IR WAV: {out_path}" in buf.getvalue() and "channels=2" in buf.getvalue()
    rate, written = wavfile.read(str(out_path))
    ref = oracle.deconvolve(np.concatenate([g["recorded"], pad]), g["sweep"])
    assert rate == SR and written.dtype == np.float32 and written.shape == ref.shape
    assert float(np.max(np.abs(written - ref))) <= 2e-4 * float(np.abs(ref).max())


def test_waterfall_db_reference_strings_and_smoothing_on_the_device():
    """Every db_reference other than "slice_max" means the global maximum (waterfall.py:328-336) - also the string
    "none", which used to select the library's internal raw mode; the log-frequency smoothing option
    (waterfall.py:140-185) runs in csd_smooth_kernel for both reference levels."""
    from audio_analysis_b200 import waterfall as wf
    from audio_analysis_b200.synth import synth_batch_numpy
    x = synth_batch_numpy(3, 96000, SR, first_index=70)
    for ref in ("global_max", "none", "whatever", "slice_max"):
        for smooth in (0, 7):
            s = wf.WaterfallAnalysisSettings(db_reference=ref, smoothing_log_bins=smooth, f_min_hz=40.0, f_max_hz=15000.0)
            got = wf.analyse_waterfall_batch(x, SR, None, s)
            p = oracle.CsdParams(db_reference=ref, smoothing_log_bins=smooth, f_min_hz=40.0, f_max_hz=15000.0)
            for i in range(x.shape[0]):
                w = oracle.waterfall(x[i], SR, p)
                assert got[i].slice_magnitude_rel_db.max() == 0.0
                _check_slices(f"dbref_{ref}/smooth{smooth}", got[i].slice_magnitude_rel_db, w["rel_db"])


def test_band_smoothing_path_beyond_65535_units():
    """edc_smoothing_window_samples > 1 in rt60bands goes through box_smooth_kernel with one grid row per (row, band)
    unit: more than 65,535 units in one call (3000 rows x 26 bands) must work (grid.y limit) and agree with a
    small batch of the same rows."""
    from audio_analysis_b200 import decay, engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_torch
    n = 4800
    x = synth_batch_torch(3000, n, SR, device="cuda", first_index=12)
    s = rb.Rt60BandsAnalysisSettings(band_mode="third",
                                     decay_settings=decay.DecayAnalysisSettings(edc_smoothing_window_samples=5))
    big = rb.run_rt60_bands(engine.make_rows(x), SR, s)[0].cpu().numpy()
    small = rb.run_rt60_bands(engine.make_rows(x[2990:]), SR, s)[0].cpu().numpy()
    assert np.array_equal(big[2990:], small, equal_nan=True)
    ob = oracle.band_metrics(x[2999].cpu().numpy(), SR, oracle.BandParams(
        band_mode="third", decay=oracle.DecayParams(edc_smoothing_window_samples=5)))
    want = np.array([[np.nan if ob["metrics"][b["name"]][0] is None else ob["metrics"][b["name"]][0]] for b in ob["bands"]])
    check_band_table("band_smoothing/65535", big[2999][:, :1], want)


def test_edc_staged_variant_is_bit_identical(monkeypatch):
    """edc_staged_kernel (csrc/edc_staged.cu, opt-in through AA_EDC_STAGED): every CTA of the row's cluster keeps its
    span of the row in shared memory between the two phases, so the row crosses HBM once.  Same arithmetic as
    edc_kernel: curves, fits and start indices must be bit-identical."""
    import torch
    from audio_analysis_b200 import decay, engine
    from audio_analysis_b200.synth import synth_batch_torch
    s = decay.DecayAnalysisSettings(compute_edt=True)
    for n in (288000, 50022, 624000):
        rows = engine.make_rows(synth_batch_torch(5, n, SR, device="cuda", first_index=3))
        monkeypatch.delenv("AA_EDC_STAGED", raising=False)
        base = decay.run_decay(rows, SR, s)
        for t in ("8", "16"):
            monkeypatch.setenv("AA_EDC_STAGED", t)
            got = decay.run_decay(rows, SR, s)
            assert torch.equal(got.start, base.start) and torch.equal(got.status, base.status)
            st = base.start.cpu().numpy()
            for r in range(5):
                assert torch.equal(got.edc_db[r, st[r]:n], base.edc_db[r, st[r]:n])
            assert torch.allclose(got.fits, base.fits, rtol=1e-12, atol=0, equal_nan=True)
    monkeypatch.delenv("AA_EDC_STAGED", raising=False)


def test_butterworth_filterbank_mode_against_scipy_sosfiltfilt():
    """filterbank="butterworth" (opt-in; BASELINE.json north_star stage 1; csrc/sos.cu: chunked 2 x 2 state-matrix
    scan with float64 carries, forward-backward).  PARITY UNPINNED against the reference - it has no such filter
     - so the pin is SciPy: band signals equal scipy.signal.sosfiltfilt of scipy.signal.butter sections,
    and the band RT60s equal the oracle's decay fits run on those SciPy-filtered signals."""
    from scipy import signal
    from audio_analysis_b200 import engine, rt60bands as rb
    from audio_analysis_b200.synth import synth_batch_numpy
    for mode, n in (("third", 96000), ("three", 60000), ("octave", 50022)):
        x = synth_batch_numpy(3, n, SR, first_index=21)
        s = rb.Rt60BandsAnalysisSettings(band_mode=mode, filterbank="butterworth", butterworth_order=4, include_t20=True)
        metrics, bands, plan, start, sig = rb.run_rt60_bands(engine.make_rows(x), SR, s, want_band_signals=True)
        metrics, sig, start = metrics.cpu().numpy(), sig.cpu().numpy(), start.cpu().numpy()
        nyq = 0.5 * SR
        for bi, b in enumerate(bands):
            lo = None if b.low_edge_hz is None else float(np.clip(b.low_edge_hz, 1.0, nyq - 1.0))
            hi = None if b.high_edge_hz is None else float(np.clip(b.high_edge_hz, 1.0, nyq - 1.0))
            if b.kind == "bandpass":
                sos = signal.butter(4, [lo, hi], btype="bandpass", fs=SR, output="sos")
            elif b.kind == "lowpass":
                sos = signal.butter(4, hi, btype="lowpass", fs=SR, output="sos")
            else:
                sos = signal.butter(4, lo, btype="highpass", fs=SR, output="sos")
            for r in range(x.shape[0]):
                want = signal.sosfiltfilt(sos, x[r].astype(np.float64))
                err = float(np.max(np.abs(sig[r, bi] - want))) / float(np.max(np.abs(want)))
                note(f"butterworth_band_signal_rel_err[{mode}]", err)
                assert err <= 2e-5, (mode, b.name, r, err)
                y = want.astype(np.float32)[start[r]:]
                t, curve, _ = oracle.ref_port.schroeder_edc_db(y, SR, oracle.DecayParams(trim_to_peak=False))
                for c, rng in enumerate(((-5.0, -35.0), (-5.0, -25.0))):
                    f = oracle.ref_port.line_fit(t, curve, rng, -80.0)
                    got = metrics[r, bi, c]
                    assert (f is None) == bool(np.isnan(got)), (mode, b.name, r, c)
                    if f is not None:
                        assert seconds_close(got, f["rt60"]), (mode, b.name, r, c, got, f["rt60"])
    # the default is the reference's filter, and the two filters are different estimators
    s0 = rb.Rt60BandsAnalysisSettings(band_mode="third")
    assert s0.filterbank == "fft_mask"
    x = synth_batch_numpy(1, 96000, SR, first_index=21)
    a = rb.run_rt60_bands(engine.make_rows(x), SR, s0)[0].cpu().numpy()
    b = rb.run_rt60_bands(engine.make_rows(x), SR, rb.Rt60BandsAnalysisSettings(band_mode="third", filterbank="butterworth"))[0].cpu().numpy()
    assert not np.allclose(a[:, :, 0], b[:, :, 0], rtol=1e-4)
    with pytest.raises(ValueError, match="Unknown filterbank"):
        rb.run_rt60_bands(engine.make_rows(x), SR, rb.Rt60BandsAnalysisSettings(filterbank="elliptic"))
