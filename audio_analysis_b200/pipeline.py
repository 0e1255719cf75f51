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
                 validate: bool = False, edc_out=None, bands_options=None) -> HotPathOut:
    ds = decay_settings or DecayAnalysisSettings()
    dec = run_decay(rows, sample_rate_hz, ds, want_edc=want_edc, edc_out=edc_out)
    metrics = bands = plan = None
    if do_bands:
        bs = bands_settings or Rt60BandsAnalysisSettings()
        bd = bs.decay_settings
        same_zero = (bool(bd.trim_to_peak) == bool(ds.trim_to_peak)
                     and float(bd.ignore_leading_seconds) == float(ds.ignore_leading_seconds))
        metrics, bands, plan, _, _ = run_rt60_bands(rows, sample_rate_hz, bs, start_idx=dec.start if same_zero else None,
                                                    options=bands_options)
    csd = csd_plan = None
    if do_csd:
        ws = waterfall_settings or WaterfallAnalysisSettings()
        peaks = dec.start if (ds.trim_to_peak and float(ds.ignore_leading_seconds) <= 0.0) else None
        csd, csd_plan = run_waterfall(rows, sample_rate_hz, ws, peaks=peaks, validate=validate)
    return HotPathOut(dec, metrics, bands, plan, csd, csd_plan)


@dataclass
class HostPathOut:
    """Result of `run_hot_path_host`: pinned host copies of the tables / CSD slices (complete after
    `torch.cuda.synchronize()` or `wait()`), device-resident EDC curves, and the byte counts moved."""

    start: torch.Tensor                 # [B] int32 (pinned host)
    fits: torch.Tensor                  # [B, 3, 8] float64
    early10: torch.Tensor               # [B] float64
    status: torch.Tensor                # [B] int32
    band_rt60: Optional[torch.Tensor]   # [B, n_bands, 3] float64
    csd: Optional[torch.Tensor]         # [B, S_max, F] float32
    edc_db: Optional[torch.Tensor]      # [B, pitch] float32, DEVICE (plot data, fetched on demand)
    bands: Optional[List[BandDefinition]]
    h2d_bytes: int
    d2h_bytes: int
    done: Optional[torch.cuda.Event] = None

    def wait(self) -> None:
        if self.done is not None:
            self.done.synchronize()


def run_hot_path_host(x_host, sample_rate_hz: int,
                      decay_settings: Optional[DecayAnalysisSettings] = None,
                      bands_settings: Optional[Rt60BandsAnalysisSettings] = None,
                      waterfall_settings: Optional[WaterfallAnalysisSettings] = None,
                      chunk_rows: int = 256, want_edc: bool = True, do_bands: bool = True, do_csd: bool = True,
                      device=None, reuse: Optional[HostPathOut] = None, bands_options=None) -> HostPathOut:
    """The hot path for a HOST batch [B, N] (numpy array or torch CPU tensor, ideally pinned): float32 samples, or
    int16 mono PCM as it sits in the reference's WAV inputs (decoded on the device with io.py:46-64's rule,
    sample / 32768; half the bytes over PCIe).

    The batch is streamed through the device in chunks of `chunk_rows` rows on three streams: the
    host-to-device copy of chunk i + 1 and the device-to-host copy of chunk i - 1's tables and CSD slices
    overlap the kernels of chunk i, so for large batches the transfer cost disappears behind the compute.
    Nothing blocks the host: call `.wait()` (or synchronise the device) before reading the host tensors.
    `reuse`: a previous result whose pinned host tensors (and device EDC buffer) are overwritten instead of
    allocating new ones - pinning a gigabyte of host memory costs more than analysing it.
    """
    import numpy as np

    dev = engine._device(device)
    xh = torch.from_numpy(x_host) if isinstance(x_host, np.ndarray) else x_host
    if xh.dim() != 2 or xh.dtype not in (torch.float32, torch.int16) or xh.is_cuda:
        raise ValueError("expected a [rows, samples] float32 or int16 host array")
    pcm = xh.dtype == torch.int16
    b, n = int(xh.shape[0]), int(xh.shape[1])
    pitch = max(4, engine.round_up4(n))
    chunk = max(1, min(int(chunk_rows), b))
    ds = decay_settings or DecayAnalysisSettings()
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs = [torch.zeros((chunk, pitch), dtype=torch.int16 if pcm else torch.float32, device=dev) for _ in range(2)]
    free_ev = [None, None]          # kernels that read buffer k have finished
    edc_all = None
    if want_edc:
        if reuse is not None and reuse.edc_db is not None and tuple(reuse.edc_db.shape) == (b, pitch):
            edc_all = reuse.edc_db
        else:
            edc_all = torch.empty((b, pitch), dtype=torch.float32, device=dev)
    host = {}
    if reuse is not None:
        reuse.wait()
        for name in ("start", "fits", "early10", "status", "band_rt60", "csd"):
            t = getattr(reuse, name)
            if t is not None and int(t.shape[0]) == b:
                host[name] = t
    bands = None
    d2h = 0
    hl = np.full(chunk, n, dtype=np.int64)

    def pinned_like(name, t, rows_total):
        want = (rows_total,) + tuple(t.shape[1:])
        if name not in host or tuple(host[name].shape) != want or host[name].dtype != t.dtype:
            host[name] = torch.empty(want, dtype=t.dtype, pin_memory=True)
        return host[name]

    n_chunks = (b + chunk - 1) // chunk
    in_ready = [None] * n_chunks

    def stage(i):
        k = i & 1
        r0, r1 = i * chunk, min(b, (i + 1) * chunk)
        with torch.cuda.stream(s_in):
            if free_ev[k] is not None:
                s_in.wait_event(free_ev[k])
            bufs[k][: r1 - r0, :n].copy_(xh[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
            in_ready[i] = ev

    stage(0)
    for i in range(n_chunks):
        if i + 1 < n_chunks:
            stage(i + 1)
        k = i & 1
        r0, r1 = i * chunk, min(b, (i + 1) * chunk)
        main.wait_event(in_ready[i])
        if pcm:
            rows = engine.pcm16_to_rows(bufs[k][: r1 - r0], n, 1, False)
        else:
            rows = engine.Rows(x=bufs[k][: r1 - r0], lengths=None, n=n, host_lengths=hl[: r1 - r0])
        smoothing = int(ds.edc_smoothing_window_samples or 0) > 1
        out = run_hot_path(rows, sample_rate_hz, ds, bands_settings, waterfall_settings, want_edc=want_edc,
                           do_bands=do_bands, do_csd=do_csd, validate=False,
                           edc_out=edc_all[r0:r1] if (want_edc and not smoothing) else None, bands_options=bands_options)
        if want_edc and smoothing:
            edc_all[r0:r1].copy_(out.decay.edc_db)
        ev = torch.cuda.Event()
        ev.record(main)
        free_ev[k] = ev
        bands = out.bands
        tables = {"start": out.decay.start, "fits": out.decay.fits, "early10": out.decay.early10,
                  "status": out.decay.status}
        if out.band_metrics is not None:
            tables["band_rt60"] = out.band_metrics
        if out.csd is not None:
            tables["csd"] = out.csd
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev)
            for name, t in tables.items():
                pinned_like(name, t, b)[r0:r1].copy_(t, non_blocking=True)
                t.record_stream(s_out)
                d2h += t.numel() * t.element_size()
    done = torch.cuda.Event()
    done.record(s_out)
    main.wait_stream(s_out)
    return HostPathOut(start=host["start"], fits=host["fits"], early10=host["early10"], status=host["status"],
                       band_rt60=host.get("band_rt60") if do_bands else None, csd=host.get("csd") if do_csd else None,
                       edc_db=edc_all, bands=bands,
                       h2d_bytes=b * n * (2 if pcm else 4), d2h_bytes=d2h, done=done)
