#!/usr/bin/env python
"""H2D bandwidth from pinned memory: alone, and while the hot path runs on another stream."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from audio_analysis_b200 import decay, engine, rt60bands, waterfall
from audio_analysis_b200.pipeline import run_hot_path
from audio_analysis_b200.synth import synth_batch_torch

dev = torch.device("cuda", 0)
n, rows = 288000, 512
xh = torch.empty((rows, n), dtype=torch.float32, pin_memory=True)
xh.normal_()
xd = torch.empty((rows, n), dtype=torch.float32, device=dev)
s_in = torch.cuda.Stream(dev)

def h2d(reps=5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_in):
        a.record(s_in)
        for _ in range(reps):
            xd.copy_(xh, non_blocking=True)
        b.record(s_in)
    return a, b, reps

a, b, r = h2d(); torch.cuda.synchronize()
a, b, r = h2d(); torch.cuda.synchronize()
print(f"H2D alone: {rows * n * 4 * r / a.elapsed_time(b) / 1e6:.1f} GB/s")
x = synth_batch_torch(rows, n, 48000, device=dev)
rws = engine.make_rows(x)
ds = decay.DecayAnalysisSettings(compute_edt=True)
bs = rt60bands.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
ws = waterfall.WaterfallAnalysisSettings()
run_hot_path(rws, 48000, ds, bs, ws); torch.cuda.synchronize()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
for _ in range(4):
    run_hot_path(rws, 48000, ds, bs, ws)
c1.record()
a, b, r = h2d(reps=4)
torch.cuda.synchronize()
print(f"compute 4 x {rows} rows: {c0.elapsed_time(c1):.1f} ms; concurrent H2D: {rows * n * 4 * r / a.elapsed_time(b) / 1e6:.1f} GB/s "
      f"({a.elapsed_time(b):.1f} ms)")
t0 = time.perf_counter()
for _ in range(4):
    run_hot_path(rws, 48000, ds, bs, ws)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time for 4 steps: {(t1 - t0) * 1e3:.1f} ms, device done after {(t2 - t0) * 1e3:.1f} ms")
