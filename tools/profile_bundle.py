#!/usr/bin/env python
"""cProfile of run_bundle_report on a synthetic config-3 bundle (256 stereo 4 s PCM16 taps): where the host time goes."""
import cProfile
import json
import pstats
import shutil
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from audio_analysis_b200 import wire  # noqa: E402
from audio_analysis_b200.bundle import BundleRunSettings, run_bundle_report  # noqa: E402
from audio_analysis_b200.report import ReportSettings  # noqa: E402
from audio_analysis_b200.rt60bands import Rt60BandsAnalysisSettings  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

SR, frames, n_taps = 48000, 192000, int(sys.argv[1]) if len(sys.argv) > 1 else 256
root = Path("/dev/shm/aa_profile_bundle")
shutil.rmtree(root, ignore_errors=True)
(root / "taps").mkdir(parents=True)
names = [f"tap_{i:03d}" for i in range(n_taps)]
for i0 in range(0, n_taps, 64):
    x = synth_batch_torch(2 * min(64, n_taps - i0), frames, SR, first_index=2 * i0, device="cuda")
    pcm = (x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16).cpu().numpy()
    for k in range(pcm.shape[0] // 2):
        wire.write_pcm16_wav(root / "taps" / f"{names[i0 + k]}.wav", np.stack([pcm[2 * k], pcm[2 * k + 1]], axis=1), SR)
(root / "meta.json").write_text(json.dumps({"sample_rate_hz": SR, "length_samples": frames, "taps": names}))
rs = ReportSettings(rt60_bands_settings=Rt60BandsAnalysisSettings(band_mode="third"), run_spectrogram=False,
                    run_waterfall=False, run_modal_cloud=False)
st = BundleRunSettings(report_settings=rs)
run_bundle_report(root, st)
t0 = time.perf_counter()
run_bundle_report(root, st)
print(f"warm call: {(time.perf_counter() - t0) * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
run_bundle_report(root, st)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
shutil.rmtree(root, ignore_errors=True)
