#!/usr/bin/env python
"""Tail of the band-RT60 change caused by decimated synthesis: decimated path against the full-rate device path
(which matches the oracle to ~4e-5, tools/decimation_error.py) over thousands of synthetic rows."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import audio_analysis_b200._native as nat  # noqa: E402
from audio_analysis_b200 import engine, rt60bands as rb  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2048)
ap.add_argument("--samples", type=int, default=288000)
ap.add_argument("--limits", type=int, nargs="+", default=[4, 8, 12, 16, 24])
a = ap.parse_args()
lib = nat.load()
SR = 48000
x = synth_batch_torch(a.rows, a.samples, SR, device="cuda", first_index=5000)
rows = engine.make_rows(x)
names = ("T30", "T20", "EDT")
for label, s, limits in (("T30 only", rb.Rt60BandsAnalysisSettings(band_mode="third"), a.limits),
                         ("T30+T20", rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True), [16]),
                         ("T30+T20+EDT", rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True), [16])):
    cols = 1 + int(s.include_t20) + int(s.include_edt)
    lib.aa_bands_set_decimation(1)
    ref = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()[:, :, :cols]
    bands = rb._build_band_definitions(s, SR)
    for d in limits:
        lib.aa_bands_set_decimation(d)
        got = rb.run_rt60_bands(rows, SR, s)[0].cpu().numpy()[:, :, :cols]
        rel = np.abs(got - ref) / ref
        bad = (rel > 5e-3) & (np.abs(got - ref) > 1e-3)
        nanmis = int((np.isnan(got) != np.isnan(ref)).sum())
        dec = rb.decimation_classes(bands, a.samples, SR, s.transition_width_octaves, settings=s)
        print(f"== {label}: limit {d}, D per band {dec}")
        for c in range(cols):
            r = rel[:, :, c][~np.isnan(rel[:, :, c])]
            print(f"   n={a.samples} {names[c]}: metrics={r.size} max={r.max():.2e} p99.99={np.quantile(r, 0.9999):.2e} "
                  f"p99.9={np.quantile(r, 0.999):.2e} p99={np.quantile(r, 0.99):.2e} median={np.median(r):.1e} "
                  f"outside(0.5% and 1 ms)={int(bad[:, :, c].sum())} nan-mismatch={nanmis}")
            print(f"      outside per band:", " ".join(f"{int(v)}" for v in bad[:, :, c].sum(axis=0)))
lib.aa_bands_set_decimation(16)
