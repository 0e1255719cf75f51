#!/usr/bin/env python
"""Fixture for BASELINE.json configs 1-2 at their real shape: a stereo file of 1,151,844 samples per channel
(plots/example/verb_report.md:6 - 2^2 * 3 * 95987, so the band filter's length-N circular FFT needs Bluestein), run
through the UNMODIFIED reference in the build container:

    python tests/golden/make_golden_verb.py

config 1: `analyse.cli rt60bands --mono`, third-octave (mono downmix 0.5 (L + R), analyse/io.py:90)
config 2: `analyse.cli decay` + `waterfall` on the stereo file (left and right separately)

The input is NOT stored (4.6 MB of noise): `verb_pcm16` below regenerates it from the seeds on any machine (numpy's
PCG64 stream is platform independent); the fixture holds the reference's outputs, the EDC curves decimated by 64.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import fits_to_rows, import_reference, none_to_nan, synth_pcm16  # noqa: E402

N = 1151844
SR = 48000
SEEDS = (201, 202)
RT60 = (2.6, 3.1)
ONSET = 24000


def verb_pcm16() -> np.ndarray:
    """[N, 2] int16: the stereo input of the fixture."""
    return np.stack([synth_pcm16(N, SR, rt, ONSET, seed) for rt, seed in zip(RT60, SEEDS)], axis=1)


def main():
    rdecay, rbands, rwf, _ = import_reference()
    pcm = verb_pcm16()
    x = (pcm.astype(np.float32) / 32768.0).astype(np.float32)     # analyse/io.py:59
    store = {"n": np.int64(N), "sr": np.int32(SR), "pcm_checksum": np.int64(int(pcm.astype(np.int64).sum())),
             "pcm_head": pcm[ONSET:ONSET + 64].copy()}
    for c, name in enumerate(("left", "right")):
        d = rdecay.analyse_decay_for_channel(x[:, c], SR, name, rdecay.DecayAnalysisSettings(compute_edt=True))
        store[f"{name}_decay_start"] = np.int64(d.analysis_start_sample_index)
        store[f"{name}_decay_len"] = np.int64(d.edc_db.size)
        store[f"{name}_decay_edc_db_dec64"] = d.edc_db[::64].copy()
        store[f"{name}_decay_early10"] = np.float64(none_to_nan(d.early_decay_10db_time_seconds))
        store[f"{name}_decay_fits"] = fits_to_rows(d.fits)
        w = rwf.analyse_waterfall_for_channel(x[:, c], SR, name, rwf.WaterfallAnalysisSettings())
        store[f"{name}_wf_start"] = np.int64(w.analysis_start_sample_index)
        store[f"{name}_wf_length"] = np.int64(w.analysis_length_samples)
        store[f"{name}_wf_times"] = w.slice_times_seconds
        store[f"{name}_wf_rel_db"] = w.slice_magnitude_rel_db
        store["wf_freq"] = w.frequency_hz
        print(name, "decay + waterfall done")
    mono = (0.5 * (x[:, 0] + x[:, 1])).astype(np.float32)        # analyse/io.py:90
    s = rbands.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    r = rbands.analyse_rt60_bands_for_channel(mono, SR, "mono", s)
    store["mono_bands_third_metrics"] = np.array(
        [[none_to_nan(r.band_metrics_by_name[b.name].rt60_t30_seconds),
          none_to_nan(r.band_metrics_by_name[b.name].rt60_t20_seconds),
          none_to_nan(r.band_metrics_by_name[b.name].edt_seconds)] for b in r.band_definitions])
    store["mono_bands_third_names"] = np.array([b.name for b in r.band_definitions])
    np.savez_compressed(HERE / "verb_1151844.npz", **store)
    print("wrote verb_1151844.npz", {k: getattr(v, "shape", None) for k, v in store.items()})


if __name__ == "__main__":
    main()
