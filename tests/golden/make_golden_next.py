#!/usr/bin/env python
"""Golden vectors for the rows SURVEY.md 8f marks "next" (N2 spectrogram, N3 modal cloud), from the UNMODIFIED
reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_next.py

Inputs are the PCM of the existing fixtures (tests/golden/<case>.npz, written by make_golden.py).  A dense
spectrogram is 2049 x T float32, so only a spread of columns is stored, plus a float64 checksum of every column
(sum over bins), which pins all frames.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import CASES, import_reference  # noqa: E402


def main():
    import_reference()
    import analyse.spectrogram as rsp

    sr = 48000
    store = {}
    for name, *_ in CASES:
        with np.load(HERE / f"{name}.npz") as z:
            pcm = z["pcm"]
        x = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
        variants = {
            "default": rsp.SpectrogramAnalysisSettings(),
            "n1024": rsp.SpectrogramAnalysisSettings(n_fft=1024, hop_length=256, use_hann_window=False,
                                                     ignore_leading_seconds=0.01, analysis_duration_seconds=0.5,
                                                     floor_db=-100.0),
            "n256": rsp.SpectrogramAnalysisSettings(n_fft=256, hop_length=200, trim_to_peak=False),
        }
        for tag, settings in variants.items():
            r = rsp.analyse_spectrogram_for_channel(x, sr, "mono", settings)
            t = r.magnitude_db.shape[1]
            cols = np.unique(np.linspace(0, t - 1, 7).round().astype(np.int64))
            key = f"{name}/{tag}"
            store[f"{key}/start"] = np.int64(r.analysis_start_sample_index)
            store[f"{key}/length"] = np.int64(r.analysis_length_samples)
            store[f"{key}/n_frames"] = np.int64(t)
            store[f"{key}/time_last"] = np.float32(r.time_seconds[-1])
            store[f"{key}/cols"] = cols
            store[f"{key}/mag_db_cols"] = r.magnitude_db[:, cols].copy()
            store[f"{key}/col_sums"] = r.magnitude_db.astype(np.float64).sum(axis=0)
            store[f"{key}/summary"] = np.array(rsp.summarise_spectrogram_results_text([r]))
        print("spectrogram", name, {k.split("/", 1)[1]: getattr(v, "shape", None) for k, v in store.items()
                                    if k.startswith(name) and k.endswith("mag_db_cols")})
    np.savez_compressed(HERE / "spectrogram.npz", **store)

    # ---- N3 modal cloud (analyse/modalcloud.py:284): the 0.3 s fixtures are short for n_fft 8192, so the cloud is
    # also taken with n_fft 2048 / hop 128 and from two longer synthetic tails
    import analyse.modalcloud as rmc
    from make_golden import synth_pcm16

    modal = {}
    inputs = {}
    for name, *_ in CASES:
        with np.load(HERE / f"{name}.npz") as z:
            inputs[name] = z["pcm"]
    inputs["long_96000"] = synth_pcm16(96000, sr, 0.8, 9000, 201)
    inputs["long_144000"] = synth_pcm16(144000, sr, 1.4, 12000, 202, floor_db=-70.0)
    modal["long_96000/pcm"] = inputs["long_96000"]
    modal["long_144000/pcm"] = inputs["long_144000"]
    for name, pcm in inputs.items():
        x = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
        variants = {
            "n2048": rmc.ModalCloudAnalysisSettings(n_fft=2048, hop_length=128),
            "n2048_t20": rmc.ModalCloudAnalysisSettings(n_fft=2048, hop_length=128, metric="t20", log_bins_per_octave=12,
                                                        ignore_leading_seconds=0.005, min_fit_points=6),
            "n1024_edt": rmc.ModalCloudAnalysisSettings(n_fft=1024, hop_length=64, metric="edt", f_min_hz=100.0,
                                                        f_max_hz=12000.0, use_hann_window=False),
        }
        if pcm.size >= 90000:
            variants["default"] = rmc.ModalCloudAnalysisSettings()
        for tag, settings in variants.items():
            r = rmc.analyse_modal_cloud_for_channel(x, sr, "mono", settings)
            key = f"{name}/{tag}"
            modal[f"{key}/start"] = np.int64(r.analysis_start_sample_index)
            modal[f"{key}/length"] = np.int64(r.analysis_length_samples)
            modal[f"{key}/metric"] = np.array(r.metric)
            modal[f"{key}/points"] = np.array([[q.centre_hz, q.rt60_seconds, q.r_squared] for q in r.points],
                                              dtype=np.float64).reshape(-1, 3)
            modal[f"{key}/summary"] = np.array(rmc.summarise_modal_cloud_results_text([r]))
            print("modalcloud", key, "points", len(r.points))
    np.savez_compressed(HERE / "modalcloud.npz", **modal)
    print("done")


if __name__ == "__main__":
    main()
