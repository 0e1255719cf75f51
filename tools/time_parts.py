#!/usr/bin/env python
"""Device time of the three analyses (CUDA events), for quick A/B runs on the GPU box."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from audio_analysis_b200 import decay, engine, modalcloud, rt60bands, spectrogram, waterfall  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=512)
ap.add_argument("--samples", type=int, default=288000)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
x = synth_batch_torch(a.rows, a.samples, 48000, device="cuda")
rows = engine.make_rows(x)
ds = decay.DecayAnalysisSettings(compute_edt=True)
bs = rt60bands.Rt60BandsAnalysisSettings(band_mode="third", decay_settings=ds)
ws = waterfall.WaterfallAnalysisSettings()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = decay.run_decay(rows, 48000, ds)
start = out.start
gb = a.rows * a.samples * 4 / 1e9
t = timed(lambda: decay.run_decay(rows, 48000, ds), a.reps * 3)
print(f"decay+edc   {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  {2 * gb / t * 1e3:8.1f} GB/s (8 B/sample)")
t = timed(lambda: decay.run_decay(rows, 48000, ds, want_edc=False), a.reps * 3)
print(f"decay only  {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  {gb / t * 1e3:8.1f} GB/s (4 B/sample)")
t = timed(lambda: rt60bands.run_rt60_bands(rows, 48000, bs, start_idx=start), a.reps)
print(f"rt60 third  {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  {t * 1e3 / a.rows / 26:6.3f} us/band-unit")
t = timed(lambda: waterfall.run_waterfall(rows, 48000, ws, peaks=start, validate=False), a.reps * 3)
print(f"csd         {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  {a.rows * 418e3 / t / 1e6:8.1f} GB/s (418 KB/row)")
if a.rows <= 256:
    ss = spectrogram.SpectrogramAnalysisSettings()
    t = timed(lambda: spectrogram.run_spectrogram(rows, 48000, ss, peaks=start), a.reps)
    frames = 1 + (a.samples - 16000 - 4096) // 512
    print(f"spectrogram {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  {a.rows * (a.samples * 4 + frames * 2049 * 4) / t / 1e6:8.1f} GB/s "
          f"(x in + ~{frames} x 2049 dB out)")
    ms = modalcloud.ModalCloudAnalysisSettings()
    t = timed(lambda: modalcloud.run_modal_cloud(rows, 48000, ms, peaks=start), a.reps)
    print(f"modal cloud {t:9.3f} ms  {t * 1e3 / a.rows:8.2f} us/row  (n_fft 8192 STFT + 240 log bins x fits)")
