#!/usr/bin/env python
"""Find the (row, band) units where the default band path differs most from the full-rate path and save those rows
(gpurun_out/cert_debug.npz) for offline analysis with the oracle."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from audio_analysis_b200 import engine, rt60bands as rb  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4096)
ap.add_argument("--samples", type=int, default=288000)
ap.add_argument("--first-index", type=int, default=9000)
ap.add_argument("--top", type=int, default=6)
a = ap.parse_args()
SR = 48000
x = synth_batch_torch(a.rows, a.samples, SR, device="cuda", first_index=a.first_index)
rows = engine.make_rows(x)
s = rb.Rt60BandsAnalysisSettings(band_mode="third")
run = lambda o: rb.run_rt60_bands(rows, SR, s, options=o)[0].cpu().numpy()[:, :, 0]
ref = run(engine.bands_options(max_decimation=1))
got = run(None)
d16 = run(engine.bands_options(refine=False))
d4 = run(engine.bands_options(max_decimation=4, refine=False))
rel = np.abs(got - ref) / ref
order = np.argsort(-np.nan_to_num(rel).ravel())[: a.top]
keep = {}
for k in order:
    r, b = divmod(int(k), rel.shape[1])
    kept = "level0" if got[r, b] == d16[r, b] else ("level1" if got[r, b] == d4[r, b] else "full")
    print(f"row {r} band {b}: full {ref[r, b]:.6f} default {got[r, b]:.6f} ({kept}) d16 {d16[r, b]:.6f} d4 {d4[r, b]:.6f} rel {rel[r, b]:.2e}")
    keep[f"row{r}"] = x[r].cpu().numpy()
out = Path("gpurun_out")
out.mkdir(exist_ok=True)
np.savez_compressed(out / "cert_debug.npz", **keep)
