"""pytest configuration: the `gpu` marker and shared fixtures."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = [
    "smooth_48000",
    "smooth_60000",
    "radix13_62400",
    "prime_factor_50022",
    "odd_45001",
    "onset_at_zero_48000",
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name: str) -> dict:
    with np.load(GOLDEN_DIR / f"{name}.npz") as z:
        d = {k: z[k] for k in z.files}
    if "pcm" in d:
        # analyse/io.py:59 - int16 PCM / 32768 -> float32
        d["x"] = (d["pcm"].astype(np.float32) / 32768.0).astype(np.float32)
    return d


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    d = load_golden(request.param)
    d["name"] = request.param
    return d


def pytest_sessionfinish(session, exitstatus):
    try:
        from _parity import dump_observed

        dump_observed()
    except Exception:
        pass
