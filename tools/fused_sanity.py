#!/usr/bin/env python
"""Two rows of each batch length through the default band path (band_fused_kernel + refinement levels): the
smallest case that exercises the fused kernel with and without unit lists.  Prints the levels, the units per level
and how many metrics are finite."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from audio_analysis_b200 import engine, rt60bands  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

for n in (288000, 192000):
    rows = engine.make_rows(synth_batch_torch(2, n, 48000, device="cuda", first_index=5))
    m = rt60bands.run_rt60_bands(rows, 48000, rt60bands.Rt60BandsAnalysisSettings(band_mode="third"))[0]
    torch.cuda.synchronize()
    print(n, "levels", engine.last_bands_stats["limits"], "units", engine.last_bands_stats["units"],
          "finite", int(torch.isfinite(m[:, :, 0]).sum()), "of", m[:, :, 0].numel())
