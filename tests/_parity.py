"""Parity helpers for the GPU tests: tolerances from BASELINE.json's north_star, and a log of the
errors actually observed (written to gpurun_out/parity_observed.json at session end)."""

from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

# north_star tolerances
EDC_TOL_DB = 0.01          # down to -60 dB
FIT_REL_TOL = 0.005        # T20/T30/RT60/EDT: 0.5 % ...
FIT_ABS_TOL_S = 0.001      # ... or 1 ms, whichever is larger
CSD_REL_TOL = 1e-4         # CSD linear magnitudes, relative
CSD_TOL_DB = 20.0 * np.log10(1.0 + CSD_REL_TOL)   # = 8.7e-4 dB

OBSERVED = {}


def note(key: str, value: float) -> None:
    value = float(value)
    if key not in OBSERVED or value > OBSERVED[key]:
        OBSERVED[key] = value


def dump_observed() -> None:
    if not OBSERVED:
        return
    out = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parent.parent)) / "gpurun_out"
    try:
        out.mkdir(parents=True, exist_ok=True)
        (out / "parity_observed.json").write_text(json.dumps(OBSERVED, indent=1, sort_keys=True) + "\n")
    except OSError:
        pass


def seconds_close(got, want) -> bool:
    """0.5 % or 1 ms, whichever is larger."""
    return abs(got - want) <= max(FIT_REL_TOL * abs(want), FIT_ABS_TOL_S)


def check_edc(tag: str, got: np.ndarray, want: np.ndarray) -> None:
    assert got.shape == want.shape, (tag, got.shape, want.shape)
    sel = want >= -60.0
    err = float(np.max(np.abs(got[sel].astype(np.float64) - want[sel].astype(np.float64)))) if sel.any() else 0.0
    note(f"edc_db_max_abs_err_to_-60dB[{tag}]", err)
    note("edc_db_max_abs_err_to_-60dB", err)
    assert err <= EDC_TOL_DB, f"{tag}: EDC differs by {err} dB above -60 dB"
    # below -60 dB the bar is not stated; keep it honest anyway (noise floor region)
    whole = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    note("edc_db_max_abs_err_full_curve", whole)
    assert whole <= 0.05, f"{tag}: EDC differs by {whole} dB somewhere on the curve"


def check_fit_rows(tag: str, got: np.ndarray, want: np.ndarray) -> None:
    """rows [3, >=7]: t_start, t_end, slope, intercept, r2, rt60, (n_points,) valid/present."""
    for r, name in enumerate(("EDT", "T20", "T30")):
        g_ok = bool(got[r][-1] == 1.0)
        w_ok = bool(want[r][-1] == 1.0)
        assert g_ok == w_ok, f"{tag}/{name}: presence {g_ok} != {w_ok}"
        if not w_ok:
            continue
        g, w = got[r], want[r]
        for c, what in ((0, "t_start"), (1, "t_end")):
            note(f"fit_{what}_abs_err_s", abs(g[c] - w[c]))
            assert abs(g[c] - w[c]) <= FIT_ABS_TOL_S, f"{tag}/{name}/{what}: {g[c]} vs {w[c]}"
        note("fit_rt60_rel_err", abs(g[5] - w[5]) / abs(w[5]))
        note("fit_slope_rel_err", abs(g[2] - w[2]) / abs(w[2]))
        assert seconds_close(g[5], w[5]), f"{tag}/{name}/rt60: {g[5]} vs {w[5]}"
        assert abs(g[2] - w[2]) <= FIT_REL_TOL * abs(w[2]), f"{tag}/{name}/slope: {g[2]} vs {w[2]}"
        note("fit_intercept_abs_err_db", abs(g[3] - w[3]))
        assert abs(g[3] - w[3]) <= 0.05, f"{tag}/{name}/intercept: {g[3]} vs {w[3]}"
        note("fit_r2_abs_err", abs(g[4] - w[4]))
        assert abs(g[4] - w[4]) <= 1e-4, f"{tag}/{name}/r2: {g[4]} vs {w[4]}"


def check_band_table(tag: str, got: np.ndarray, want: np.ndarray, strict: bool = True) -> None:
    """[n_bands, 3] RT60 seconds, NaN = None."""
    assert got.shape == want.shape, (tag, got.shape, want.shape)
    bad = []
    for i in range(want.shape[0]):
        for j in range(want.shape[1]):
            g, w = got[i, j], want[i, j]
            if np.isnan(w) or np.isnan(g):
                if np.isnan(w) != np.isnan(g):
                    bad.append((i, j, g, w))
                continue
            note("band_rt60_rel_err", abs(g - w) / abs(w))
            if not seconds_close(g, w):
                bad.append((i, j, g, w))
    if strict:
        assert not bad, f"{tag}: band metrics outside 0.5 %/1 ms: {bad[:6]}"
    return bad
