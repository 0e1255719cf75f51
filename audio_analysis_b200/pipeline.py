"""The whole hot path for a device batch: decay -> third-octave (or other) rt60bands -> CSD.

One peak search serves all three analyses (the decay kernel finds it), results stay on the
device, and nothing synchronises with the host between the launches.  Used by the batched
bundle path, `bench.py` and the multi-GPU driver.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import engine
from .decay import DecayAnalysisSettings, run_decay
from .rt60bands import BandDefinition, Rt60BandsAnalysisSettings, run_rt60_bands
from .waterfall import WaterfallAnalysisSettings, WaterfallPlan, run_waterfall


@dataclass
class HotPathOut:
    decay: engine.DecayOut
    band_metrics: Optional[torch.Tensor]      # [B, n_bands, 3] float64 RT60 seconds (NaN = None)
    bands: Optional[List[BandDefinition]]
    band_fit_plan: Optional[list]
    csd: Optional[torch.Tensor]               # [B, S_max, F] float32 relative dB
    csd_plan: Optional[WaterfallPlan]


def run_hot_path(rows: engine.Rows, sample_rate_hz: int,
                 decay_settings: Optional[DecayAnalysisSettings] = None,
                 bands_settings: Optional[Rt60BandsAnalysisSettings] = None,
                 waterfall_settings: Optional[WaterfallAnalysisSettings] = None,
                 want_edc: bool = True, do_bands: bool = True, do_csd: bool = True,
                 validate: bool = False) -> HotPathOut:
    ds = decay_settings or DecayAnalysisSettings()
    dec = run_decay(rows, sample_rate_hz, ds, want_edc=want_edc)
    metrics = bands = plan = None
    if do_bands:
        bs = bands_settings or Rt60BandsAnalysisSettings()
        bd = bs.decay_settings
        same_zero = (bool(bd.trim_to_peak) == bool(ds.trim_to_peak)
                     and float(bd.ignore_leading_seconds) == float(ds.ignore_leading_seconds))
        metrics, bands, plan, _, _ = run_rt60_bands(rows, sample_rate_hz, bs, start_idx=dec.start if same_zero else None)
    csd = csd_plan = None
    if do_csd:
        ws = waterfall_settings or WaterfallAnalysisSettings()
        peaks = dec.start if (ds.trim_to_peak and float(ds.ignore_leading_seconds) <= 0.0) else None
        csd, csd_plan = run_waterfall(rows, sample_rate_hz, ws, peaks=peaks, validate=validate)
    return HotPathOut(dec, metrics, bands, plan, csd, csd_plan)
