#!/usr/bin/env python
"""EDC kernel variants side by side: base (row read twice) against staged (row kept in the cluster's shared memory),
bit-equality of their outputs, device time per 8192 x 288,000 batch."""
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

CHILD = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, ".")
from audio_analysis_b200 import decay, engine
from audio_analysis_b200.synth import synth_batch_torch
rows_n, n = int(sys.argv[1]), int(sys.argv[2])
x = synth_batch_torch(rows_n, n, 48000, device="cuda")
rows = engine.make_rows(x)
s = decay.DecayAnalysisSettings(compute_edt=True)
out = decay.run_decay(rows, 48000, s)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    out = decay.run_decay(rows, 48000, s)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
h = torch.zeros(1, dtype=torch.float64, device="cuda")
chk = float(out.edc_db.double().nan_to_num().sum()) + float(out.fits.nan_to_num().sum())
print(f"{os.environ.get('AA_EDC_STAGED', 'default'):>8} rows {rows_n} n {n}: {ms:.3f} ms  {8.0 * rows_n * n / ms / 1e6:.0f} GB/s algorithmic  checksum {chk!r}")
'''
for rows_n, n in ((8192, 288000), (8192, 192000)):
    for mode in ("0", "8", "16"):
        env = dict(os.environ)
        if mode is None:
            env.pop("AA_EDC_STAGED", None)
        else:
            env["AA_EDC_STAGED"] = mode
        subprocess.run([sys.executable, "-c", CHILD, str(rows_n), str(n)], env=env, check=False)
