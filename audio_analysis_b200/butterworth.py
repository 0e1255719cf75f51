"""Butterworth band filters as second-order sections, for the OPT-IN `filterbank="butterworth"` mode of rt60bands.

The reference has no such filter (its band filter is the circular FFT mask of analyse/rt60bands.py:116-175; SURVEY.md
D1): this mode exists because BASELINE.json's north_star names it, and it is pinned to SciPy instead -
`scipy.signal.butter(order, ..., output="sos")` + `scipy.signal.sosfiltfilt` - by tests/test_gpu_parity.py.  The
design below is the textbook chain SciPy uses (analog prototype -> frequency transform with pre-warped edges ->
bilinear transform), done in numpy so that the product does not import scipy.signal; the pairing of poles into
sections is free (a cascade commutes), so each conjugate pole pair simply takes one of the zero pairs.
The filtering itself runs on the device (csrc/sos.cu).
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def _prototype_poles(order: int) -> np.ndarray:
    k = np.arange(-order + 1, order, 2)
    return -np.exp(1j * np.pi * k / (2 * order))       # left half plane, unit circle


def butter_sos(order: int, kind: str, low_hz: Optional[float], high_hz: Optional[float], sample_rate_hz: float) -> np.ndarray:
    """[n_sections, 6] float64 sections (b0 b1 b2 1 a1 a2) of a Butterworth low / high / band pass.
    kind: "lowpass" (edge high_hz), "highpass" (edge low_hz), "bandpass" (low_hz .. high_hz; 2 * order poles)."""
    if order < 1:
        raise ValueError("order must be >= 1")
    nyq = 0.5 * float(sample_rate_hz)
    fs = 2.0
    warp = lambda f: 2.0 * fs * np.tan(np.pi * (float(f) / nyq) / fs)
    p = _prototype_poles(order)
    if kind == "lowpass":
        if not 0.0 < high_hz < nyq:
            raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
        wo = warp(high_hz)
        poles_s, zeros_s, gain = p * wo, np.zeros(0, complex), wo ** order
    elif kind == "highpass":
        if not 0.0 < low_hz < nyq:
            raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
        wo = warp(low_hz)
        poles_s, zeros_s, gain = wo / p, np.zeros(order, complex), 1.0     # prototype gain 1, prod(-p) = 1
    elif kind == "bandpass":
        if not 0.0 < low_hz < high_hz < nyq:
            raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
        w1, w2 = warp(low_hz), warp(high_hz)
        bw, wo = w2 - w1, np.sqrt(w1 * w2)
        half = p * bw / 2.0
        root = np.sqrt(half * half - wo * wo + 0j)
        poles_s = np.concatenate([half + root, half - root])
        zeros_s, gain = np.zeros(order, complex), bw ** order
    else:
        raise ValueError(f"Unknown band kind: {kind}")
    # bilinear transform at fs = 2
    fs2 = 2.0 * fs
    poles_z = (fs2 + poles_s) / (fs2 - poles_s)
    zeros_z = (fs2 + zeros_s) / (fs2 - zeros_s)
    gain = gain * np.real(np.prod(fs2 - zeros_s) / np.prod(fs2 - poles_s))
    zeros_z = np.concatenate([zeros_z, -np.ones(poles_z.size - zeros_z.size)])    # zeros at infinity -> z = -1
    # sections: conjugate pole pairs (and one real pole for odd low / high pass orders)
    cplx = poles_z[np.imag(poles_z) > 1e-14]
    real = np.real(poles_z[np.abs(np.imag(poles_z)) <= 1e-14])
    plus = int(np.sum(np.real(zeros_z) > 0))      # zeros at +1
    minus = zeros_z.size - plus                  # zeros at -1
    sections = []
    for pole in sorted(cplx, key=lambda c: abs(c)):
        a = [1.0, -2.0 * float(np.real(pole)), float(abs(pole) ** 2)]
        if kind == "bandpass":
            b = [1.0, 0.0, -1.0]                  # one zero at +1, one at -1
            plus, minus = plus - 1, minus - 1
        elif plus >= 2:
            b = [1.0, -2.0, 1.0]
            plus -= 2
        else:
            b = [1.0, 2.0, 1.0]
            minus -= 2
        sections.append(b + a)
    for pole in real:                            # first-order section
        if plus >= 1:
            b, plus = [1.0, -1.0, 0.0], plus - 1
        else:
            b, minus = [1.0, 1.0, 0.0], minus - 1
        sections.append(b + [1.0, -float(pole), 0.0])
    sos = np.array(sections, dtype=np.float64)
    # spread the gain over the sections (keeps every section's output near the signal's scale)
    g = abs(gain) ** (1.0 / sos.shape[0])
    sos[:, :3] *= g
    if gain < 0:
        sos[0, :3] *= -1.0
    return sos


def sosfilt_zi(sos: np.ndarray) -> np.ndarray:
    """[n_sections, 2]: state of every section (transposed direct form II) in the steady state of a unit step input -
    what scipy.signal.sosfilt_zi returns; sosfiltfilt scales it by the first sample of each pass."""
    sos = np.asarray(sos, dtype=np.float64)
    zi = np.zeros((sos.shape[0], 2))
    level = 1.0                                   # steady input level of the section
    for k, (b0, b1, b2, _, a1, a2) in enumerate(sos):
        out = level * (b0 + b1 + b2) / (1.0 + a1 + a2)
        zi[k, 0] = out - b0 * level               # y = b0 x + z0
        zi[k, 1] = b2 * level - a2 * out          # z1 = b2 x - a2 y
        level = out
    return zi


def filtfilt_padlen(sos: np.ndarray) -> int:
    """scipy.signal.sosfiltfilt's default odd-extension length: 3 * (2 n_sections + 1 - min(#b2 == 0, #a2 == 0))."""
    sos = np.asarray(sos)
    ntaps = 2 * sos.shape[0] + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    return 3 * ntaps


def band_sections(bands: Sequence[Tuple[str, Optional[float], Optional[float]]], order: int, sample_rate_hz: float):
    """(sos [n_bands, n_sections, 6], zi [n_bands, n_sections, 2], padlen [n_bands]) for (kind, low, high) triples.
    Low / high passes have half the sections of a band pass of the same order: they are padded with unit sections."""
    all_sos = [butter_sos(order, kind, lo, hi, sample_rate_hz) for kind, lo, hi in bands]
    edges = np.array([filtfilt_padlen(s) for s in all_sos], dtype=np.int32)
    n_sec = max(s.shape[0] for s in all_sos)
    for i, s in enumerate(all_sos):               # pad low / high passes with unit sections next to band passes
        if s.shape[0] < n_sec:
            pad = np.tile(np.array([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]), (n_sec - s.shape[0], 1))
            all_sos[i] = np.vstack([s, pad])
    sos = np.stack(all_sos)
    zi = np.stack([sosfilt_zi(s) for s in all_sos])
    return sos, zi, edges
