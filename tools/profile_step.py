#!/usr/bin/env python
"""One or two hot-path steps over a small batch, for ncu (launch list / --set full captures).

    python tools/profile_step.py --rows 96 --steps 2
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from audio_analysis_b200 import decay, engine, rt60bands, waterfall  # noqa: E402
from audio_analysis_b200.pipeline import run_hot_path  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=96)
ap.add_argument("--samples", type=int, default=288000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--only", choices=["all", "decay", "bands", "csd"], default="all")
a = ap.parse_args()

x = synth_batch_torch(a.rows, a.samples, 48000, device="cuda")
rows = engine.make_rows(x)
ds = decay.DecayAnalysisSettings(compute_edt=True)
bs = rt60bands.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
ws = waterfall.WaterfallAnalysisSettings()
torch.cuda.synchronize()
for _ in range(a.steps):
    if a.only == "all":
        run_hot_path(rows, 48000, ds, bs, ws)
    elif a.only == "decay":
        decay.run_decay(rows, 48000, ds)
    elif a.only == "bands":
        rt60bands.run_rt60_bands(rows, 48000, bs)
    else:
        waterfall.run_waterfall(rows, 48000, ws)
    torch.cuda.synchronize()
print("done")
