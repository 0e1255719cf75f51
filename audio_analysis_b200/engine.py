"""Batched device engine: torch tensors as buffers, the C ABI for all arithmetic.

Everything here is plumbing (allocation, host<->device copies, stream handles,
plan / workspace caches).  The analysis itself happens in
libaudio_analysis_b200.so; there is no torch arithmetic and no CPU path.

Rows = analysed channels (the reference's 1-D `samples`).  A `Rows` object is a
device tensor [B, stride] of float32 with per-row lengths; stride % 4 == 0 so
every row starts 16-byte aligned.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as nat

ArrayLike = Union[np.ndarray, torch.Tensor, Sequence[np.ndarray]]


def _device(device=None) -> torch.device:
    nat.require_device()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _pitch(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def round_up4(n: int) -> int:
    return (int(n) + 3) & ~3


@dataclass
class Rows:
    """Device-resident batch of channels."""

    x: torch.Tensor                  # [B, stride] float32, cuda
    lengths: Optional[torch.Tensor]  # [B] int32 cuda, or None when every row has `n` samples
    n: int                           # maximum row length
    host_lengths: np.ndarray         # [B] int64 (host copy)

    @property
    def n_rows(self) -> int:
        return int(self.x.shape[0])

    @property
    def stride(self) -> int:
        # torch reports arbitrary strides for size-1 dimensions: a single row's pitch is its width
        return int(self.x.stride(0)) if self.x.shape[0] > 1 else int(self.x.shape[1])

    @property
    def uniform(self) -> bool:
        return self.lengths is None


def make_rows(data: ArrayLike, device=None, lengths: Optional[Sequence[int]] = None) -> Rows:
    """Stage channels on the device.

    Accepts a 2-D float32 numpy array / torch tensor [B, N] (CPU, ideally pinned, or
    CUDA) or a list of 1-D arrays of different lengths (zero padded to a common pitch).
    Inputs are never modified.
    """
    dev = _device(device)
    if isinstance(data, torch.Tensor):
        t = data
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2:
            raise ValueError("expected [rows, samples]")
        if t.dtype != torch.float32:
            t = t.float()
        b, n = int(t.shape[0]), int(t.shape[1])
        if t.is_cuda:
            dev = t.device
            ok = t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.stride(0) >= n and t.data_ptr() % 16 == 0
            if ok:
                x = t
            else:
                x = torch.zeros((b, max(4, round_up4(n))), dtype=torch.float32, device=dev)
                x[:, :n].copy_(t)
        else:
            stride = max(4, round_up4(n))
            if stride == n and t.is_contiguous():
                x = t.to(dev, non_blocking=True)
            else:
                x = torch.zeros((b, stride), dtype=torch.float32, device=dev)
                x[:, :n].copy_(t, non_blocking=True)
        hl = np.full(b, n, dtype=np.int64) if lengths is None else np.asarray(lengths, dtype=np.int64)
        return _finish_rows(x, n, hl, dev)
    if isinstance(data, np.ndarray) and data.ndim == 2:
        arr = np.ascontiguousarray(data, dtype=np.float32)
        b, n = arr.shape
        stride = max(4, round_up4(n))
        if stride == n:
            x = torch.from_numpy(arr).to(dev, non_blocking=False)
        else:
            x = torch.zeros((b, stride), dtype=torch.float32, device=dev)
            x[:, :n].copy_(torch.from_numpy(arr))
        hl = np.full(b, n, dtype=np.int64) if lengths is None else np.asarray(lengths, dtype=np.int64)
        return _finish_rows(x, n, hl, dev)
    # ragged list of 1-D arrays
    chans = [np.asarray(c, dtype=np.float32) for c in data]
    for c in chans:
        if c.ndim != 1:
            raise ValueError("expects 1D mono arrays")
    b = len(chans)
    n = max((c.size for c in chans), default=0)
    stride = max(4, round_up4(n))
    host = torch.zeros((b, stride), dtype=torch.float32)
    hv = host.numpy()
    for i, c in enumerate(chans):
        hv[i, : c.size] = c
    x = host.to(dev)
    hl = np.array([c.size for c in chans], dtype=np.int64)
    return _finish_rows(x, n, hl, dev)


def _finish_rows(x: torch.Tensor, n: int, hl: np.ndarray, dev: torch.device) -> Rows:
    if n > nat.AA_MAX_ROW_SAMPLES:
        raise ValueError(f"rows longer than {nat.AA_MAX_ROW_SAMPLES} samples are not supported")
    if hl.size and int(hl.max(initial=0)) > n:
        raise ValueError("a row length exceeds the buffer width")
    uniform = bool(hl.size == 0 or np.all(hl == n))
    lengths = None if uniform else torch.from_numpy(hl.astype(np.int32)).to(x.device)
    return Rows(x=x, lengths=lengths, n=int(n), host_lengths=hl)


# --------------------------------------------------------------------------------------
# decay
# --------------------------------------------------------------------------------------


@dataclass
class DecayOut:
    start: torch.Tensor          # [B] int32
    fits: torch.Tensor           # [B, 3, 8] float64 (rows follow `fit_ranges` order)
    early10: torch.Tensor        # [B] float64, NaN = None
    status: torch.Tensor         # [B] int32
    edc_db: Optional[torch.Tensor]  # [B, stride] float32, entry j = curve at absolute sample j (valid j >= start)


def decay_params(sample_rate_hz: int, trim_to_peak: bool, ignore_leading_seconds: float, edc_floor_db: float,
                 edc_epsilon: float, fit_lower_limit_db: float, fit_ranges: Sequence[Tuple[float, float]],
                 want_early10: bool = True) -> nat.DecayParams:
    if len(fit_ranges) > nat.AA_MAX_FITS:
        raise ValueError("at most 3 fit ranges")
    p = nat.DecayParams()
    p.sample_rate_hz = int(sample_rate_hz)
    p.trim_to_peak = 1 if trim_to_peak else 0
    ign = 0
    if ignore_leading_seconds > 0.0:  # analyse/decay.py:140-141
        ign = max(0, int(round(float(ignore_leading_seconds) * float(sample_rate_hz))))
    p.ignore_leading_samples = min(ign, nat.AA_MAX_ROW_SAMPLES)
    p.smoothing_window = 0
    p.edc_floor_db = float(edc_floor_db)
    p.edc_epsilon = float(edc_epsilon)
    p.fit_lower_limit_db = float(fit_lower_limit_db)
    p.n_fits = len(fit_ranges)
    p.want_early10 = 1 if want_early10 else 0
    for i, (hi, lo) in enumerate(fit_ranges):
        hi, lo = float(hi), float(lo)
        if lo > hi:  # analyse/decay.py:216-217
            raise ValueError("range_db should be (higher_db, lower_db), e.g. (-5, -25).")
        p.fit_hi_db[i] = hi
        p.fit_lo_db[i] = lo
    return p


def decay(rows: Rows, params: nat.DecayParams, want_edc: bool = True,
          start_in: Optional[torch.Tensor] = None, edc_out: Optional[torch.Tensor] = None) -> DecayOut:
    """`edc_out`: optional [B, stride] float32 device buffer (same pitch as the rows) to receive the curves."""
    lib = nat.load()
    dev = rows.x.device
    b = rows.n_rows
    start = torch.empty(b, dtype=torch.int32, device=dev)
    fits = torch.empty((b, nat.AA_MAX_FITS, nat.AA_FIT_COLS), dtype=torch.float64, device=dev)
    early = torch.empty(b, dtype=torch.float64, device=dev)
    status = torch.empty(b, dtype=torch.int32, device=dev)
    edc = None
    if want_edc:
        if edc_out is not None:
            if (edc_out.dtype != torch.float32 or not edc_out.is_cuda or tuple(edc_out.shape) != (b, rows.stride)
                    or edc_out.stride(1) != 1 or (b > 1 and edc_out.stride(0) != rows.stride)):
                raise ValueError("edc_out must be a float32 device tensor [rows, stride] with the rows' pitch")
            edc = edc_out
        else:
            edc = torch.empty((b, rows.stride), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.aa_decay_batch(
            _ptr(rows.x), rows.stride, _ptr(rows.lengths), rows.n, b, C.byref(params), _ptr(start_in), _ptr(edc),
            rows.stride if want_edc else 0, _ptr(start), _ptr(fits), _ptr(early), _ptr(status), _stream_ptr()))
    return DecayOut(start=start, fits=fits, early10=early, status=status, edc_db=edc)


def curve_fits(curve: torch.Tensor, lengths: Optional[torch.Tensor], uniform_len: int,
               offsets: Optional[torch.Tensor], params: nat.DecayParams) -> Tuple[torch.Tensor, torch.Tensor]:
    """Crossings + fits on caller-supplied dB curves [B, stride] -> (fits [B,3,8], early10 [B])."""
    lib = nat.load()
    dev = curve.device
    b = int(curve.shape[0])
    fits = torch.empty((b, nat.AA_MAX_FITS, nat.AA_FIT_COLS), dtype=torch.float64, device=dev)
    early = torch.empty(b, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.aa_curve_fits_batch(_ptr(curve), _pitch(curve), _ptr(lengths), int(uniform_len),
                                          _ptr(offsets), b, C.byref(params), _ptr(fits), _ptr(early), _stream_ptr()))
    return fits, early


def box_smooth_db(curve: torch.Tensor, lengths: Optional[torch.Tensor], uniform_len: int,
                  offsets: Optional[torch.Tensor], window: int, floor_db: float) -> torch.Tensor:
    """np.convolve(curve, ones(w)/w, "same") then the floor, per row (analyse/decay.py:161-167)."""
    lib = nat.load()
    dev = curve.device
    out = torch.empty_like(curve)
    with torch.cuda.device(dev):
        nat.check(lib.aa_box_smooth_db_batch(_ptr(curve), _ptr(out), _pitch(curve), _ptr(lengths),
                                             int(uniform_len), _ptr(offsets), int(curve.shape[0]), int(window),
                                             float(floor_db), _stream_ptr()))
    return out


def decay_smoothed(rows: Rows, params: nat.DecayParams, window: int) -> DecayOut:
    """decay() with edc_smoothing_window_samples > 1: raw dB curve -> box filter -> generic fits."""
    raw = nat.DecayParams()
    C.memmove(C.byref(raw), C.byref(params), C.sizeof(nat.DecayParams))
    raw.n_fits = 0
    raw.want_early10 = 0
    raw.edc_floor_db = float("-inf")
    first = decay(rows, raw, want_edc=True)
    smooth = box_smooth_db(first.edc_db, rows.lengths, rows.n, first.start, window, params.edc_floor_db)
    fits, early = curve_fits(smooth, rows.lengths, rows.n, first.start, params)
    return DecayOut(start=first.start, fits=fits, early10=early, status=first.status, edc_db=smooth)


def peak_index(rows: Rows) -> torch.Tensor:
    lib = nat.load()
    dev = rows.x.device
    out = torch.empty(rows.n_rows, dtype=torch.int32, device=dev)
    lengths = rows.lengths
    if lengths is None:
        lengths = torch.full((rows.n_rows,), rows.n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.aa_peak_index_batch(_ptr(rows.x), rows.stride, _ptr(lengths), rows.n_rows, _ptr(out),
                                          _stream_ptr()))
    return out


# --------------------------------------------------------------------------------------
# rt60bands
# --------------------------------------------------------------------------------------


class _BandsPlan:
    def __init__(self, n: int):
        lib = nat.load()
        h = C.c_void_p()
        nat.check(lib.aa_bands_plan_create(int(n), C.byref(h)))
        self.handle = h
        self.n = int(n)
        self.kind = int(lib.aa_bands_plan_kind(h))

    def close(self):
        if self.handle:
            nat.load().aa_bands_plan_destroy(self.handle)
            self.handle = None


_bands_plans: Dict[Tuple[int, int], _BandsPlan] = {}
_workspaces: Dict[int, torch.Tensor] = {}
DEFAULT_WORKSPACE_BYTES = 40 << 30   # of 180 GB: the bench's 8192 rows in ONE chunk.  Every chunk pays the refinement levels'
# small launches and stream synchronisations again: 12 GB (3 chunks) 55.3 ms per step, 24 GB 54.6, 40 GB 53.8


def slice_rows(rows: Rows, r0: int, r1: int) -> Rows:
    """Rows [r0, r1) of a batch as a batch of their own (views, no copy)."""
    lengths = None if rows.lengths is None else rows.lengths[r0:r1].contiguous()
    return Rows(x=rows.x[r0:r1], lengths=lengths, n=rows.n, host_lengths=rows.host_lengths[r0:r1])


def bands_plan(n: int, dev: torch.device) -> _BandsPlan:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), int(n))
    plan = _bands_plans.get(key)
    if plan is None:
        with torch.cuda.device(dev):
            plan = _BandsPlan(n)
        _bands_plans[key] = plan
    return plan


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces[key] = None  # release before growing
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def release_caches() -> None:
    for p in _bands_plans.values():
        p.close()
    _bands_plans.clear()
    for p in _deconv_plans.values():
        p.close()
    _deconv_plans.clear()
    _workspaces.clear()


def band_array(bands: Sequence[Tuple[int, Optional[float], Optional[float]]]):
    """(kind, low_edge, high_edge) triples -> ctypes aa_band array."""
    if not 1 <= len(bands) <= nat.AA_MAX_BANDS:
        raise ValueError(f"between 1 and {nat.AA_MAX_BANDS} bands")
    arr = (nat.Band * len(bands))()
    for i, (kind, lo, hi) in enumerate(bands):
        arr[i].kind = int(kind)
        arr[i].low_edge_hz = 0.0 if lo is None else float(lo)
        arr[i].high_edge_hz = 0.0 if hi is None else float(hi)
    return arr


def bands_classify(n: int, dev: torch.device, bands: Sequence[Tuple[int, Optional[float], Optional[float]]],
                   transition_width_octaves: float, sample_rate_hz: int,
                   params: Optional[nat.DecayParams] = None) -> List[int]:
    """Decimation factor per band for metrics-only calls on rows of `n` samples (aa_bands_classify); `params`
    carries the requested fits (short / early ranges such as EDT keep the decimation at <= 4)."""
    lib = nat.load()
    plan = bands_plan(int(n), dev)
    out = (C.c_int32 * len(bands))()
    nat.check(lib.aa_bands_classify(plan.handle, band_array(bands), len(bands), float(transition_width_octaves),
                                    int(sample_rate_hz), C.byref(params) if params is not None else None, out))
    return [int(v) for v in out]


# units per refinement level of the most recent rt60bands call on this process: {"limits": [...], "units": [...]}
last_bands_stats: Dict[str, object] = {"limits": [], "units": [], "chunks": 0}


def bands_options(max_decimation: int = 0, refine: Optional[bool] = None, cert_safety: float = 0.0,
                  cert_window: float = 0.0, cert_min_span: int = 0, cert_tol: float = 0.0) -> nat.BandsOptions:
    """Per-call options of the band stage (aa_bands_options).  max_decimation 1 = every band at full rate (the
    reference's estimator), 0 = process default; refine False keeps every first-level result uncertified."""
    o = nat.BandsOptions()
    o.max_decimation = int(max_decimation)
    o.refine = 0 if refine is None else (1 if refine else -1)
    o.cert_safety = float(cert_safety)
    o.cert_window = float(cert_window)
    o.cert_min_span = int(cert_min_span)
    o.cert_tol = float(cert_tol)
    return o


def rt60bands(rows: Rows, start_idx: torch.Tensor, bands: Sequence[Tuple[int, Optional[float], Optional[float]]],
              transition_width_octaves: float, params: nat.DecayParams, want_band_signals: bool = False,
              workspace_bytes: Optional[int] = None, options: Optional[nat.BandsOptions] = None):
    """RT60 table [B, n_bands, 3] (float64, NaN = None; columns follow params' fit order)."""
    if not rows.uniform:
        raise ValueError("rt60bands needs rows of equal length (group channels by length)")
    lib = nat.load()
    dev = rows.x.device
    b, nb = rows.n_rows, len(bands)
    plan = bands_plan(rows.n, dev)
    one = int(lib.aa_bands_workspace_bytes(plan.handle, 1, nb))
    budget = int(workspace_bytes or os.environ.get("AA_BANDS_WORKSPACE_BYTES") or DEFAULT_WORKSPACE_BYTES)
    if budget < one:
        budget = one
    allrows = int(lib.aa_bands_workspace_bytes(plan.handle, b, nb))
    ws_bytes = min(budget, allrows)
    ws = workspace(ws_bytes, dev)
    metrics = torch.empty((b, nb, nat.AA_MAX_FITS), dtype=torch.float64, device=dev)
    sig = torch.empty((b, nb, rows.n), dtype=torch.float32, device=dev) if want_band_signals else None
    barr = band_array(bands)
    stats = nat.BandsStats()
    with torch.cuda.device(dev):
        nat.check(lib.aa_rt60bands_batch_ex(
            plan.handle, _ptr(rows.x), rows.stride, b, _ptr(start_idx), barr, nb, float(transition_width_octaves),
            C.byref(params), _ptr(metrics), _ptr(sig), _ptr(ws), ws_bytes, _stream_ptr(),
            C.byref(options) if options is not None else None, C.byref(stats)))
    last_bands_stats["limits"] = [int(stats.level_limit[i]) for i in range(int(stats.n_levels))]
    last_bands_stats["units"] = [int(stats.level_units[i]) for i in range(int(stats.n_levels))]
    last_bands_stats["chunks"] = int(stats.n_chunks)
    return metrics, sig


def rt60bands_sos(rows: Rows, start_idx: torch.Tensor, sos: np.ndarray, zi: np.ndarray, edges: np.ndarray,
                  params: nat.DecayParams, want_band_signals: bool = False, workspace_bytes: Optional[int] = None):
    """RT60 table [B, n_bands, 3] with the opt-in Butterworth filterbank (aa_rt60bands_sos_batch): `sos`
    [n_bands, n_sections, 6], `zi` [n_bands, n_sections, 2], `edges` [n_bands] from butterworth.band_sections."""
    if not rows.uniform:
        raise ValueError("rt60bands needs rows of equal length (group channels by length)")
    lib = nat.load()
    dev = rows.x.device
    b = rows.n_rows
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    zi = np.ascontiguousarray(zi, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.int32)
    nb, n_sec = int(sos.shape[0]), int(sos.shape[1])
    budget = int(workspace_bytes or os.environ.get("AA_BANDS_WORKSPACE_BYTES") or DEFAULT_WORKSPACE_BYTES)
    one = int(lib.aa_sos_workspace_bytes(rows.n, n_sec, 1, nb))
    allrows = int(lib.aa_sos_workspace_bytes(rows.n, n_sec, b, nb))
    ws_bytes = min(max(budget, one), allrows)
    ws = workspace(ws_bytes, dev)
    metrics = torch.empty((b, nb, nat.AA_MAX_FITS), dtype=torch.float64, device=dev)
    sig = torch.empty((b, nb, rows.n), dtype=torch.float32, device=dev) if want_band_signals else None
    with torch.cuda.device(dev):
        nat.check(lib.aa_rt60bands_sos_batch(
            _ptr(rows.x), rows.stride, rows.n, b, _ptr(start_idx), sos.ctypes.data, zi.ctypes.data, edges.ctypes.data,
            nb, n_sec, C.byref(params), _ptr(metrics), _ptr(sig), _ptr(ws), ws_bytes, _stream_ptr()))
    return metrics, sig


# --------------------------------------------------------------------------------------
# waterfall / CSD
# --------------------------------------------------------------------------------------


def csd_params(n_fft: int, hop_length: int, use_hann_window: bool, bin_lo: int, n_bins: int, db_reference: str,
               floor_db: float, dynamic_range_db: float, raw_db: bool = False) -> nat.CsdParams:
    """`db_reference`: the user's setting - "slice_max", anything else means the global maximum like the reference
    (waterfall.py:328-336).  `raw_db`: internal - no reference level and no clip (spectrogram-style output, and the
    input of the smoothing kernel)."""
    p = nat.CsdParams()
    p.n_fft = int(n_fft)
    p.hop_length = int(hop_length)
    p.use_hann_window = 1 if use_hann_window else 0
    p.bin_lo = int(bin_lo)
    p.n_bins = int(n_bins)
    if raw_db:
        p.db_reference = nat.AA_DBREF_NONE
    else:
        p.db_reference = nat.AA_DBREF_SLICE_MAX if str(db_reference).lower() == "slice_max" else nat.AA_DBREF_GLOBAL_MAX
    p.floor_db = float(floor_db)
    p.dynamic_range_db = float(max(10.0, dynamic_range_db))  # analyse/waterfall.py:338
    return p


def csd(rows: Rows, start_idx: torch.Tensor, frame_idx: torch.Tensor, n_slices: torch.Tensor,
        params: nat.CsdParams) -> torch.Tensor:
    """Relative-dB slices [B, max_slices, n_bins] float32."""
    lib = nat.load()
    dev = rows.x.device
    b = rows.n_rows
    max_slices = int(frame_idx.shape[1])
    out = torch.empty((b, max_slices, int(params.n_bins)), dtype=torch.float32, device=dev)   # unused slices: kernel
    with torch.cuda.device(dev):
        nat.check(lib.aa_csd_batch(_ptr(rows.x), rows.stride, _ptr(rows.lengths), rows.n, b, _ptr(start_idx),
                                   _ptr(frame_idx), _ptr(n_slices), max_slices, C.byref(params), _ptr(out),
                                   _stream_ptr()))
    return out


def csd_plan(rows: Rows, peaks: Optional[torch.Tensor], ignore_samples: int, keep_samples: int, n_fft: int, hop_length: int,
             table: torch.Tensor, counts: torch.Tensor):
    """(start [B], span [B], frames [B, S_max], n_slices [B]) int32 on the device (aa_csd_plan_batch)."""
    lib = nat.load()
    dev = rows.x.device
    b = rows.n_rows
    width = int(table.shape[1])
    start = torch.empty(b, dtype=torch.int32, device=dev)
    span = torch.empty(b, dtype=torch.int32, device=dev)
    frames = torch.empty((b, width), dtype=torch.int32, device=dev)
    nslices = torch.empty(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.aa_csd_plan_batch(_ptr(rows.lengths), rows.n, b, _ptr(peaks), int(ignore_samples), int(keep_samples),
                                        int(n_fft), int(hop_length), _ptr(table), _ptr(counts), int(table.shape[0]) - 1,
                                        width, _ptr(start), _ptr(span), _ptr(frames), _ptr(nslices), _stream_ptr()))
    return start, span, frames, nslices


def csd_smooth(slices: torch.Tensor, n_slices: torch.Tensor, k0: int, f_in: np.ndarray, grid: np.ndarray,
               grid_seg: np.ndarray, bin_seg: np.ndarray, window: int, db_reference: str,
               dynamic_range_db: float) -> torch.Tensor:
    """In place: raw dB slices [B, S_max, F] -> log-frequency smoothed, referenced and clipped (aa_csd_smooth_batch)."""
    lib = nat.load()
    dev = slices.device
    b, s_max, f = (int(v) for v in slices.shape)
    tf = torch.from_numpy(np.ascontiguousarray(f_in, dtype=np.float64)).to(dev)
    tg = torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float64)).to(dev)
    tgs = torch.from_numpy(np.ascontiguousarray(grid_seg, dtype=np.int32)).to(dev)
    tbs = torch.from_numpy(np.ascontiguousarray(bin_seg, dtype=np.int32)).to(dev)
    ref = nat.AA_DBREF_SLICE_MAX if str(db_reference).lower() == "slice_max" else nat.AA_DBREF_GLOBAL_MAX
    with torch.cuda.device(dev):
        nat.check(lib.aa_csd_smooth_batch(_ptr(slices), b, _ptr(n_slices), s_max, f, int(k0), int(f_in.size), _ptr(tf),
                                          _ptr(tg), int(grid.size), _ptr(tgs), _ptr(tbs), int(window), ref,
                                          float(max(10.0, dynamic_range_db)), _stream_ptr()))
    return slices


# --------------------------------------------------------------------------------------
# PCM16 decode and deconvolution
# --------------------------------------------------------------------------------------


def stft_db(rows: Rows, start_idx: torch.Tensor, span: torch.Tensor, n_fft: int, hop_length: int,
            use_hann_window: bool, floor_db: float, frames_max: int) -> torch.Tensor:
    """Dense STFT magnitude in dB, [B, frames_max, n_fft/2 + 1] float32 (frame-major; frames a row does not have
    stay zero).  n_fft in {1024, 2048, 4096, 8192} (aa_stft_db_batch)."""
    lib = nat.load()
    dev = rows.x.device
    b = rows.n_rows
    out = torch.zeros((b, int(frames_max), int(n_fft) // 2 + 1), dtype=torch.float32, device=dev)
    if b == 0 or frames_max <= 0:
        return out
    with torch.cuda.device(dev):
        for r0 in range(0, b, 65535):   # grid.y limit
            r1 = min(b, r0 + 65535)
            x = rows.x[r0:r1]
            nat.check(lib.aa_stft_db_batch(
                _ptr(x), rows.stride, _ptr(rows.lengths[r0:r1]) if rows.lengths is not None else None, rows.n, r1 - r0,
                _ptr(start_idx[r0:r1]), _ptr(span[r0:r1]), int(n_fft), int(hop_length), 1 if use_hann_window else 0,
                float(floor_db), int(frames_max), _ptr(out[r0:r1]), _stream_ptr()))
    return out


def modal_cloud(mag_db: torch.Tensor, n_frames: torch.Tensor, bin_ranges: torch.Tensor, hop_length: int,
                sample_rate_hz: int, range_db: Tuple[float, float], fit_lower_limit_db: float, min_fit_points: int,
                floor_db: float, min_peak_db_above_floor: float) -> torch.Tensor:
    """[B, n_bins, 2] float64 (rt60 seconds, r^2; NaN = no point) from a dense STFT [B, T_max, F] (aa_modal_cloud_batch)."""
    lib = nat.load()
    dev = mag_db.device
    b, t_max, f = (int(v) for v in mag_db.shape)
    n_bins = int(bin_ranges.shape[0])
    out = torch.full((b, n_bins, 2), float("nan"), dtype=torch.float64, device=dev)
    if b == 0 or t_max == 0:
        return out
    ws = torch.empty((b * n_bins * t_max,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.aa_modal_cloud_batch(
            _ptr(mag_db), b, t_max, f, _ptr(n_frames), _ptr(bin_ranges), n_bins, int(hop_length), int(sample_rate_hz),
            float(range_db[0]), float(range_db[1]), float(fit_lower_limit_db), int(min_fit_points), float(floor_db),
            float(min_peak_db_above_floor), _ptr(ws), _ptr(out), _stream_ptr()))
    return out


def pcm16_to_rows(pcm: torch.Tensor, n_frames: int, n_channels: int, mono_downmix: bool) -> Rows:
    """pcm: [files, n_frames*n_channels] int16 on the device -> float32 rows (io.py:46-95)."""
    lib = nat.load()
    dev = pcm.device
    files = int(pcm.shape[0])
    per_file = 1 if (mono_downmix or n_channels == 1) else n_channels
    stride = round_up4(n_frames)
    out = torch.empty((files * per_file, stride), dtype=torch.float32, device=dev)
    if stride > n_frames:
        out[:, n_frames:].zero_()          # row padding (< 4 samples): nothing else needs clearing
    with torch.cuda.device(dev):
        nat.check(lib.aa_pcm16_to_float_batch(_ptr(pcm), _pitch(pcm), files, int(n_frames), int(n_channels),
                                              1 if (mono_downmix and n_channels == 2) else 0, _ptr(out), stride,
                                              _stream_ptr()))
    hl = np.full(files * per_file, n_frames, dtype=np.int64)
    return _finish_rows(out, int(n_frames), hl, dev)


class _DeconvPlan:
    def __init__(self, n_fft: int):
        lib = nat.load()
        h = C.c_void_p()
        nat.check(lib.aa_deconv_plan_create(int(n_fft), C.byref(h)))
        self.handle = h
        self.n_fft = int(n_fft)

    def close(self):
        if self.handle:
            nat.load().aa_deconv_plan_destroy(self.handle)
            self.handle = None


_deconv_plans: Dict[Tuple[int, int], _DeconvPlan] = {}


def deconvolve(rec: Rows, sweep: torch.Tensor, n_fft: int, regularization_relative: float, remove_dc: bool,
               target_peak: float, norm_group: int, n_out: int) -> torch.Tensor:
    """Impulse responses [B, round_up4(n_out)] float32 (analyse/deconvolve.py:124-193)."""
    if not rec.uniform:
        raise ValueError("deconvolve needs recordings of equal length")
    lib = nat.load()
    dev = rec.x.device
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), int(n_fft))
    plan = _deconv_plans.get(key)
    if plan is None:
        with torch.cuda.device(dev):
            plan = _DeconvPlan(n_fft)
        _deconv_plans[key] = plan
    b = rec.n_rows
    ws = workspace(int(lib.aa_deconv_workspace_bytes(plan.handle, b)), dev)
    stride = round_up4(n_out)
    out = torch.zeros((b, stride), dtype=torch.float32, device=dev)
    sweep = sweep.to(dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        nat.check(lib.aa_deconvolve_batch(plan.handle, _ptr(sweep), int(sweep.numel()), _ptr(rec.x), rec.stride,
                                          rec.n, b, float(regularization_relative), 1 if remove_dc else 0,
                                          float(target_peak), int(norm_group), int(n_out), _ptr(out), stride,
                                          _ptr(ws), ws.numel(), _stream_ptr()))
    return out
