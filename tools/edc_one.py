#!/usr/bin/env python
"""A few launches of the decay kernel on a resident batch (for ncu): rows, samples from argv."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from audio_analysis_b200 import decay, engine  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

rows_n, n = int(sys.argv[1]), int(sys.argv[2])
rows = engine.make_rows(synth_batch_torch(rows_n, n, 48000, device="cuda"))
s = decay.DecayAnalysisSettings(compute_edt=True)
for _ in range(3):
    decay.run_decay(rows, 48000, s)
torch.cuda.synchronize()
print("ok")
