"""The oracle (oracle/ref_port.py) against outputs of the unmodified reference.

Fixtures: tests/golden/*.npz, produced by tests/golden/make_golden.py which
imports /root/reference/analyse/* in the build container.  These tests pin the
oracle; the GPU parity tests then compare the CUDA path with the oracle.
"""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from conftest import load_golden, load_verb_case

SR = 48000


def rows_from_fits(fits: dict) -> np.ndarray:
    out = np.full((3, 7), np.nan)
    for r, name in enumerate(("EDT", "T20", "T30")):
        f = fits.get(name)
        if f is None:
            out[r, 6] = 0.0
        else:
            out[r] = [f["t_start"], f["t_end"], f["slope"], f["intercept"], f["r2"], f["rt60"], 1.0]
    return out


def assert_rows_equal(got, want, rtol=1e-9):
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=rtol, atol=1e-12)


def test_decay_matches_reference(golden):
    x = golden["x"]
    d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True))
    assert d["start"] == int(golden["decay_start"])
    assert np.array_equal(d["edc_db"], golden["decay_edc_db"])  # bit-exact float32 curve
    assert_rows_equal(rows_from_fits(d["fits"]), golden["decay_fits"])
    want = float(golden["decay_early10"])
    assert (d["early10"] is None) == np.isnan(want)
    if d["early10"] is not None:
        assert d["early10"] == pytest.approx(want, rel=1e-12)


def test_decay_variants_match_reference(golden):
    x = golden["x"]
    d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True, ignore_leading_seconds=0.01))
    assert d["start"] == int(golden["decay_ign_start"])
    assert_rows_equal(rows_from_fits(d["fits"]), golden["decay_ign_fits"])
    d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True, edc_smoothing_window_samples=31))
    assert_rows_equal(rows_from_fits(d["fits"]), golden["decay_smooth_fits"])
    assert np.array_equal(d["edc_db"][::16], golden["decay_smooth_edc_db_dec16"])
    d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True, trim_to_peak=False))
    assert d["start"] == 0
    assert_rows_equal(rows_from_fits(d["fits"]), golden["decay_notrim_fits"])


@pytest.mark.parametrize("mode", ["three", "octave", "third"])
def test_band_metrics_match_reference(golden, mode):
    p = oracle.BandParams(band_mode=mode, include_t20=True, include_edt=True)
    r = oracle.band_metrics(golden["x"], SR, p)
    names = [b["name"] for b in r["bands"]]
    assert names == [str(s) for s in golden[f"bands_{mode}_names"]]
    edges = np.array([[np.nan if b["low"] is None else b["low"], np.nan if b["high"] is None else b["high"],
                       b["centre"]] for b in r["bands"]])
    want_edges = golden[f"bands_{mode}_edges"]
    assert np.array_equal(np.isnan(edges), np.isnan(want_edges))
    assert np.array_equal(edges[~np.isnan(edges)], want_edges[~np.isnan(want_edges)])
    got = np.array([[np.nan if v is None else v for v in r["metrics"][n]] for n in names])
    assert_rows_equal(got, golden[f"bands_{mode}_metrics"])


def test_band_signal_matches_reference(golden):
    x = golden["x"]
    n = x.size
    freqs = np.fft.rfftfreq(n, d=1.0 / SR).astype(np.float32)
    bands = {b["name"]: b for b in oracle.band_table(oracle.BandParams(band_mode="third"), SR)}
    for probe in ("1000Hz", "62Hz"):
        m = oracle.band_mask(freqs, bands[probe], 1.0 / 6.0, 0.5 * SR)
        y = oracle.band_filter(x, m)
        assert np.array_equal(y[::8], golden[f"band_{probe}_signal_dec8"])
        nz = np.array([int(np.argmax(m > 0)), int(m.size - np.argmax(m[::-1] > 0))])
        assert np.array_equal(nz, golden[f"band_{probe}_mask_nonzero"])


def test_waterfall_matches_reference(golden):
    x = golden["x"]
    for only_sel in (True, False):
        w = oracle.waterfall(x, SR, oracle.CsdParams(), only_selected_frames=only_sel)
        assert w["start"] == int(golden["wf_start"])
        assert w["length"] == int(golden["wf_length"])
        assert np.array_equal(w["slice_times"], golden["wf_times"])
        assert np.array_equal(w["freq"], golden["wf_freq"])
        assert np.array_equal(w["rel_db"], golden["wf_rel_db"])
    w2 = oracle.waterfall(x, SR, oracle.CsdParams(db_reference="slice_max", num_slices=7,
                                                   slice_mode="uniform_frames", dynamic_range_db=60.0))
    assert np.array_equal(w2["slice_times"], golden["wf2_times"])
    assert np.array_equal(w2["rel_db"], golden["wf2_rel_db"])
    w3 = oracle.waterfall(x, SR, oracle.CsdParams(smoothing_log_bins=9, slice_mode="uniform_time",
                                                   slice_spacing_seconds=0.08))
    assert np.array_equal(w3["slice_times"], golden["wf3_times"])
    assert np.array_equal(w3["rel_db"], golden["wf3_rel_db"])


def test_verb_sized_stereo_file_matches_reference():
    """Configs 1-2 of BASELINE.json at their real shape: 1,151,844 samples x 2 channels (2^2 * 3 * 95987)."""
    g = load_verb_case()
    for c, name in enumerate(("left", "right")):
        x = g["x"][:, c]
        d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True))
        assert d["start"] == int(g[f"{name}_decay_start"]) and d["edc_db"].size == int(g[f"{name}_decay_len"])
        assert np.array_equal(d["edc_db"][::64], g[f"{name}_decay_edc_db_dec64"])
        assert_rows_equal(rows_from_fits(d["fits"]), g[f"{name}_decay_fits"])
        assert d["early10"] == pytest.approx(float(g[f"{name}_decay_early10"]), rel=1e-12)
        w = oracle.waterfall(x, SR, oracle.CsdParams())
        assert w["start"] == int(g[f"{name}_wf_start"]) and w["length"] == int(g[f"{name}_wf_length"])
        assert np.array_equal(w["slice_times"], g[f"{name}_wf_times"]) and np.array_equal(w["freq"], g["wf_freq"])
        assert np.array_equal(w["rel_db"], g[f"{name}_wf_rel_db"])
    r = oracle.band_metrics(g["mono"], SR, oracle.BandParams(band_mode="third", include_t20=True, include_edt=True))
    names = [b["name"] for b in r["bands"]]
    assert names == [str(s) for s in g["mono_bands_third_names"]]
    got = np.array([[np.nan if v is None else v for v in r["metrics"][n]] for n in names])
    assert_rows_equal(got, g["mono_bands_third_metrics"])


def test_analytic_exponential_known_answers():
    g = load_golden("analytic")
    for rt in (0.3, 1.2):
        n = int(SR * 2.0)
        x = (10.0 ** (-3.0 * np.arange(n) / (SR * rt))).astype(np.float32)
        d = oracle.decay_metrics(x, SR, oracle.DecayParams(compute_edt=True))
        assert_rows_equal(rows_from_fits(d["fits"]), g[f"rt{rt}_fits"])
        for name in ("EDT", "T20", "T30"):
            assert d["fits"][name]["rt60"] == pytest.approx(rt, rel=2e-5)
            assert d["fits"][name]["r2"] == pytest.approx(1.0, abs=1e-9)
        assert d["early10"] == pytest.approx(rt / 6.0, rel=2e-4)
        assert d["early10"] == pytest.approx(float(g[f"rt{rt}_early10"]), rel=1e-12)
        assert np.float32(d["edc_db"][-1]) == g[f"rt{rt}_edc_last"]


def test_band_table_known_answers():
    """SURVEY.md section 4 item 3."""
    names = [b["name"] for b in oracle.band_table(oracle.BandParams(band_mode="third"), SR)]
    assert len(names) == 26 and names[0] == "39Hz" and names[-1] == "12699Hz"  # no 16 kHz band
    octv = [b["name"] for b in oracle.band_table(oracle.BandParams(band_mode="octave"), SR)]
    assert octv == ["62Hz", "125Hz", "250Hz", "500Hz", "1000Hz", "2000Hz", "4000Hz", "8000Hz", "16000Hz"]
    three = oracle.band_table(oracle.BandParams(), SR)
    assert [b["kind"] for b in three] == ["lowpass", "bandpass", "highpass"]
    assert three[1]["low"] == pytest.approx(500.0) and three[1]["high"] == pytest.approx(2000.0)


def test_deconvolve_matches_reference():
    g = load_golden("deconvolve")
    h = oracle.deconvolve(g["recorded"], g["sweep"])
    assert np.array_equal(h, g["ir"])


def test_error_cases():
    with pytest.raises(ValueError):
        oracle.schroeder_edc_db(np.zeros((4, 2), np.float32), SR, oracle.DecayParams())
    with pytest.raises(ValueError):
        oracle.schroeder_edc_db(np.array([0, 0, 1.0], np.float32), SR, oracle.DecayParams())
    with pytest.raises(ValueError):
        oracle.band_metrics(np.ones(7, np.float32), SR, oracle.BandParams())
    with pytest.raises(ValueError):
        oracle.waterfall(np.ones(4000, np.float32), SR, oracle.CsdParams())
    with pytest.raises(ValueError):
        oracle.band_table(oracle.BandParams(band_mode="quarter"), SR)
    # all-zero IR: flat 0 dB curve, no fits (decay.py:154 clamps to epsilon)
    d = oracle.decay_metrics(np.zeros(1000, np.float32), SR, oracle.DecayParams(compute_edt=True))
    assert np.all(d["edc_db"] == 0.0) and d["fits"] == {} and d["early10"] is None


# ---- spectrogram (fixtures from tests/golden/make_golden_next.py) ----------------------------

SPEC_VARIANTS = {
    "default": oracle.SpectrogramParams(),
    "n1024": oracle.SpectrogramParams(n_fft=1024, hop_length=256, use_hann_window=False, ignore_leading_seconds=0.01,
                                      analysis_duration_seconds=0.5, floor_db=-100.0),
    "n256": oracle.SpectrogramParams(n_fft=256, hop_length=200, trim_to_peak=False),
}


@pytest.mark.parametrize("variant", sorted(SPEC_VARIANTS))
def test_spectrogram_matches_reference(golden, variant):
    from conftest import GOLDEN_DIR

    with np.load(GOLDEN_DIR / "spectrogram.npz") as z:
        g = {k.split("/", 2)[2]: z[k] for k in z.files if k.startswith(f"{golden['name']}/{variant}/")}
    r = oracle.spectrogram(golden["x"], SR, SPEC_VARIANTS[variant])
    assert r["start"] == int(g["start"]) and r["length"] == int(g["length"])
    assert r["mag_db"].shape[1] == int(g["n_frames"]) and r["times"][-1] == g["time_last"]
    assert np.array_equal(r["mag_db"][:, g["cols"]], g["mag_db_cols"])          # bit-exact float32 columns
    assert np.array_equal(r["mag_db"].astype(np.float64).sum(axis=0), g["col_sums"])   # every frame pinned


# ---- modal cloud ------------------------------------------------------------------------------

MODAL_VARIANTS = {
    "default": oracle.ModalParams(),
    "n2048": oracle.ModalParams(n_fft=2048, hop_length=128),
    "n2048_t20": oracle.ModalParams(n_fft=2048, hop_length=128, metric="t20", log_bins_per_octave=12,
                                    ignore_leading_seconds=0.005, min_fit_points=6),
    "n1024_edt": oracle.ModalParams(n_fft=1024, hop_length=64, metric="edt", f_min_hz=100.0, f_max_hz=12000.0,
                                    use_hann_window=False),
}


def modal_golden():
    from conftest import GOLDEN_DIR

    with np.load(GOLDEN_DIR / "modalcloud.npz") as z:
        return {k: z[k] for k in z.files}


def modal_input(g: dict, case: str) -> np.ndarray:
    pcm = g[f"{case}/pcm"] if f"{case}/pcm" in g else load_golden(case)["pcm"]
    return (pcm.astype(np.float32) / 32768.0).astype(np.float32)


@pytest.mark.parametrize("case,variant", [("smooth_48000", "n2048"), ("odd_45001", "n2048_t20"),
                                          ("prime_factor_50022", "n1024_edt"), ("onset_at_zero_48000", "n2048"),
                                          ("long_96000", "default"), ("long_144000", "default"),
                                          ("long_144000", "n2048_t20"), ("radix13_62400", "n1024_edt")])
def test_modal_cloud_matches_reference(case, variant):
    g = modal_golden()
    r = oracle.modal_cloud(modal_input(g, case), SR, MODAL_VARIANTS[variant])
    key = f"{case}/{variant}"
    assert r["start"] == int(g[f"{key}/start"]) and r["length"] == int(g[f"{key}/length"])
    assert r["metric"] == str(g[f"{key}/metric"])
    got = np.array(r["points"], dtype=np.float64).reshape(-1, 3)
    assert np.array_equal(got, g[f"{key}/points"])       # same bins, bit-exact RT60 and r^2
