"""Bundle runner: a folder with `meta.json` and `taps/<name>.wav` (written by the C++
AnalysisRecorder, include/analysis/recorder.hpp:102-126 of the reference).

Same entry point, arguments and output layout as the reference's `analyse/bundle.py`:
`<bundle_root>/<reports_subdir>/<tap>/<tap>_report.md` per tap plus the index
`<reports_subdir>/bundle_report.md`.  The reference analyses taps one after the other and
reloads each WAV for every block (bundle.py:56-67); here all taps are loaded once and
every channel of every tap shares the same kernel launches.
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional

from .report import ReportSettings, run_reports_batch


@dataclass(frozen=True)
class BundleRunSettings:
    reports_subdir: str = "reports"
    report_settings: Optional[ReportSettings] = None


def run_bundle_report(bundle_root: str | Path, settings: Optional[BundleRunSettings] = None) -> Path:
    """Reference bundle.py:35-74.  All taps go through `report.run_reports_batch`: their PCM16 payloads are staged
    once in pinned memory and decoded on the device, and every channel shares the kernel launches.  Under
    `torch.distributed` (one process per GPU) the taps are sharded over the ranks in contiguous blocks
    (`distributed.shard_range`); every rank writes the reports of its taps, rank 0 writes the index."""
    from . import distributed

    settings = settings or BundleRunSettings()
    root = Path(bundle_root)
    meta = json.loads((root / "meta.json").read_text())
    taps: List[str] = list(meta.get("taps", []))
    reports_root = root / settings.reports_subdir
    reports_root.mkdir(parents=True, exist_ok=True)

    rank, size = distributed.world()
    lo, hi = distributed.shard_range(len(taps), rank, size)
    mine = taps[lo:hi]
    wavs = [root / "taps" / f"{tap}.wav" for tap in mine]
    bases = [reports_root / tap / tap for tap in mine]
    for b in bases:
        b.parent.mkdir(parents=True, exist_ok=True)
    run_reports_batch(wavs, bases, settings.report_settings)
    if size > 1:
        import torch.distributed as dist

        dist.barrier()

    index = reports_root / "bundle_report.md"
    if rank == 0:
        lines = [
            "# IR Bundle Report\n",
            f"**Bundle:** `{root}`\n",
            f"**Sample rate:** {meta.get('sample_rate_hz')}\n",
            f"**Length (samples):** {meta.get('length_samples')}\n",
            "\n## Taps\n",
        ]
        lines += [f"- [{tap}]({settings.reports_subdir}/{tap}/{tap}_report.md)" for tap in taps]
        index.write_text("\n".join(lines) + "\n")
    return index
