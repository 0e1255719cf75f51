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


def load_verb_case() -> dict:
    """BASELINE.json configs 1-2 at their real shape (tests/golden/make_golden_verb.py): the fixture holds the
    reference's outputs, the 1,151,844 x 2 int16 input is regenerated from its seeds and checked against the stored
    checksum.  Returns the fixture plus x [N, 2] float32 (analyse/io.py:59) and mono = 0.5 (L + R) (io.py:90)."""
    if str(GOLDEN_DIR) not in sys.path:
        sys.path.insert(0, str(GOLDEN_DIR))
    from make_golden_verb import verb_pcm16

    with np.load(GOLDEN_DIR / "verb_1151844.npz") as z:
        d = {k: z[k] for k in z.files}
    pcm = verb_pcm16()
    assert int(pcm.astype(np.int64).sum()) == int(d["pcm_checksum"]) and np.array_equal(pcm[24000:24064], d["pcm_head"])
    d["x"] = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
    d["mono"] = (0.5 * (d["x"][:, 0] + d["x"][:, 1])).astype(np.float32)
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
