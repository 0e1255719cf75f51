#!/usr/bin/env python
"""Tail of the band-RT60 difference between the default band path (decimated synthesis with certified refinement)
and the full-rate device path (max_decimation = 1: the reference's estimator, within ~4e-5 of the oracle) over
thousands of synthetic rows.  Prints, per setting, the units each refinement level processed, quantiles of the
relative difference and the number of metrics outside the 0.5 % / 1 ms tolerance (must be 0)."""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from audio_analysis_b200 import engine, rt60bands as rb  # noqa: E402
from audio_analysis_b200.synth import synth_batch_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4096)
ap.add_argument("--samples", type=int, nargs="+", default=[288000, 192000])
ap.add_argument("--first-index", type=int, default=5000)
ap.add_argument("--sweep", action="store_true", help="also try other certification settings / no refinement")
ap.add_argument("--fits", choices=["t30", "all"], default="all")
ap.add_argument("--json", type=str, default=None)
a = ap.parse_args()
SR = 48000
names = ("T30", "T20", "EDT")
report = []


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


for n in a.samples:
    x = synth_batch_torch(a.rows, n, SR, device="cuda", first_index=a.first_index)
    rows = engine.make_rows(x)
    cases = [("T30 only", rb.Rt60BandsAnalysisSettings(band_mode="third"))]
    if a.fits == "all":
        cases.append(("T30+T20+EDT", rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)))
    for label, s in cases:
        cols = 1 + int(s.include_t20) + int(s.include_edt)
        run = lambda o: rb.run_rt60_bands(rows, SR, s, options=o)[0]
        run(engine.bands_options(max_decimation=1))
        ref_t, ref_ms = timed(lambda: run(engine.bands_options(max_decimation=1)))
        ref = ref_t.cpu().numpy()[:, :, :cols]
        settings = [("default", engine.bands_options())]
        if a.sweep:
            settings += [("no refinement", engine.bands_options(refine=False)),
                         ("safety 6", engine.bands_options(cert_safety=6.0)),
                         ("safety 6 tol 3e-3", engine.bands_options(cert_safety=6.0, cert_tol=3e-3)),
                         ("tol 3e-3", engine.bands_options(cert_tol=3e-3)),
                         ("tol 5e-3", engine.bands_options(cert_tol=5e-3)),
                         ("limit 4", engine.bands_options(max_decimation=4))]
        for tag, o in settings:
            run(o)
            got_t, ms = timed(lambda: run(o))
            stats = dict(engine.last_bands_stats)
            got = got_t.cpu().numpy()[:, :, :cols]
            rel = np.abs(got - ref) / ref
            bad = (rel > 5e-3) & (np.abs(got - ref) > 1e-3)
            nanmis = int((np.isnan(got) != np.isnan(ref)).sum())
            units = a.rows * got.shape[1]
            print(f"== n={n} {label} [{tag}]: {ms:.1f} ms (full rate {ref_ms:.1f} ms)  levels {stats['limits']} units {stats['units']}"
                  f" redo fraction {sum(stats['units'][1:]) / units:.4f}")
            for c in range(cols):
                r = rel[:, :, c][~np.isnan(rel[:, :, c])]
                line = dict(n=n, fits=label, setting=tag, metric=names[c], metrics=int(r.size), max=float(r.max()),
                            p9999=float(np.quantile(r, 0.9999)), p999=float(np.quantile(r, 0.999)),
                            median=float(np.median(r)), outside=int(bad[:, :, c].sum()), nan_mismatch=nanmis,
                            ms=ms, full_rate_ms=ref_ms, levels=stats["limits"], level_units=stats["units"])
                report.append(line)
                print(f"   {names[c]}: metrics={r.size} max={r.max():.2e} p99.99={line['p9999']:.2e} p99.9={line['p999']:.2e} "
                      f"median={line['median']:.1e} outside(0.5% and 1 ms)={line['outside']} nan-mismatch={nanmis}")
                print("      outside per band:", " ".join(f"{int(v)}" for v in bad[:, :, c].sum(axis=0)),
                      "| worst per band:", " ".join(f"{v:.0e}" for v in np.nanmax(rel[:, :, c], axis=0)))
    del x, rows
    torch.cuda.empty_cache()
if a.json:
    Path(a.json).write_text(json.dumps(report, indent=1))
