#!/usr/bin/env python
"""Timing of pipeline.run_hot_path_host variants (diagnostic)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from audio_analysis_b200 import decay, engine, rt60bands, waterfall
from audio_analysis_b200.pipeline import run_hot_path, run_hot_path_host
from audio_analysis_b200.synth import synth_batch_torch

dev = torch.device("cuda", 0)
n, rows = 288000, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = synth_batch_torch(rows, n, 48000, device=dev)
xh = torch.empty((rows, n), dtype=torch.float32, pin_memory=True)
xh.copy_(x); torch.cuda.synchronize()
ds = decay.DecayAnalysisSettings(compute_edt=True)
bs = rt60bands.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
ws = waterfall.WaterfallAnalysisSettings()
rws = engine.make_rows(x)

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t2 - t0) / reps * 1e3, (t1 - t0) / reps * 1e3

print("resident            %.1f ms (host enqueue %.1f)" % timed(lambda: run_hot_path(rws, 48000, ds, bs, ws)))
for chunk in (256, 512, 1024):
    print(f"host chunk {chunk:5d}    %.1f ms (host enqueue %.1f)" % timed(lambda: run_hot_path_host(xh, 48000, ds, bs, ws, chunk_rows=chunk)))
print("host no edc         %.1f ms (host enqueue %.1f)" % timed(lambda: run_hot_path_host(xh, 48000, ds, bs, ws, chunk_rows=512, want_edc=False)))
print("host no csd         %.1f ms (host enqueue %.1f)" % timed(lambda: run_hot_path_host(xh, 48000, ds, bs, ws, chunk_rows=512, do_csd=False)))
print("host no bands       %.1f ms (host enqueue %.1f)" % timed(lambda: run_hot_path_host(xh, 48000, ds, bs, ws, chunk_rows=512, do_bands=False)))
