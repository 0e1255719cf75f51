"""Synthetic impulse responses for tests and benchmarks.

x[n] = 0 for n < n0;  g[n] * 10^(-3 (n - n0) / (sr * RT60)) for n >= n0, g ~ N(0, 1);
plus a stationary floor 10^(-90/20) * g'[n]; scaled to max|x| = 0.95.
RT60 ~ U(0.3, 2.0) s (capped so -60 dB is reached inside the buffer), onset ~ U(0.25, 0.5) s:
the leading silence keeps the zero-phase circular band filters from wrapping their
pre-ringing into the tail.  Row i uses seed `base_seed + i`.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def ir_parameters(index: int, n: int, sr: int, base_seed: int = 1234) -> Tuple[float, int]:
    rng = np.random.default_rng(base_seed + int(index))
    rt60 = float(rng.uniform(0.3, 2.0))
    onset = int(round(float(rng.uniform(0.25, 0.5)) * sr))
    onset = min(onset, max(0, n // 4))
    longest = max(0.05, 0.9 * (n - onset) / sr)  # -60 dB must fit in the remaining buffer
    return min(rt60, longest), onset


def synth_ir_numpy(index: int, n: int, sr: int = 48000, base_seed: int = 1234, floor_db: float = -90.0,
                   rt60: Optional[float] = None, onset: Optional[int] = None) -> np.ndarray:
    """One float32 IR row on the host (deterministic in (base_seed, index))."""
    auto_rt, auto_on = ir_parameters(index, n, sr, base_seed)
    rt60 = auto_rt if rt60 is None else float(rt60)
    onset = auto_on if onset is None else int(onset)
    rng = np.random.default_rng((base_seed + int(index)) * 7919 + 1)
    x = np.zeros(n, dtype=np.float64)
    t = np.arange(n - onset, dtype=np.float64) / sr
    x[onset:] = rng.standard_normal(n - onset) * 10.0 ** (-3.0 * t / rt60)
    x += 10.0 ** (floor_db / 20.0) * rng.standard_normal(n)
    x *= 0.95 / np.max(np.abs(x))
    return x.astype(np.float32)


def synth_batch_numpy(count: int, n: int, sr: int = 48000, base_seed: int = 1234, first_index: int = 0) -> np.ndarray:
    return np.stack([synth_ir_numpy(first_index + i, n, sr, base_seed) for i in range(count)], axis=0)


def synth_batch_torch(count: int, n: int, sr: int = 48000, base_seed: int = 1234, first_index: int = 0,
                      device="cuda", floor_db: float = -90.0, chunk: int = 1024):
    """[count, n] float32 on `device`, same distribution as the numpy generator (different random
    stream: torch Philox per row).  Rows can be copied back and handed to the CPU oracle, which is
    how the benchmark keeps GPU and CPU inputs identical."""
    import torch

    out = torch.empty((count, n), dtype=torch.float32, device=device)
    t_axis = torch.arange(n, dtype=torch.float32, device=device)
    for c0 in range(0, count, chunk):
        c1 = min(count, c0 + chunk)
        pars = [ir_parameters(first_index + i, n, sr, base_seed) for i in range(c0, c1)]
        rt = torch.tensor([p[0] for p in pars], dtype=torch.float32, device=device)[:, None]
        on = torch.tensor([p[1] for p in pars], dtype=torch.float32, device=device)[:, None]
        gen = torch.Generator(device=device)
        gen.manual_seed(int(base_seed) * 1000003 + int(first_index) + c0)
        g = torch.randn((c1 - c0, n), generator=gen, device=device, dtype=torch.float32)
        rel = (t_axis[None, :] - on) / float(sr)
        env = torch.where(rel >= 0, torch.pow(10.0, -3.0 * rel / rt), torch.zeros_like(rel))
        x = g * env
        g2 = torch.randn((c1 - c0, n), generator=gen, device=device, dtype=torch.float32)
        x += (10.0 ** (floor_db / 20.0)) * g2
        x *= 0.95 / x.abs().amax(dim=1, keepdim=True)
        out[c0:c1] = x
    return out
