"""Figure helpers.  matplotlib is optional and imported lazily: metrics, tables and
plot DATA (the arrays in the result dataclasses) never need it.  When it is missing,
figure rendering is skipped with a one-line note on stderr instead of failing the
analysis (the reference imports matplotlib at module import time).
"""

from __future__ import annotations

import sys
from pathlib import Path
from typing import Optional, Tuple

DEFAULT_FIGURE_SIZE = (10.0, 6.0)
DEFAULT_DPI = 100
DEFAULT_GRID = True

_warned = False


def pyplot(optional: bool = True):
    """Return matplotlib.pyplot, or None when it is not installed (and optional)."""
    global _warned
    if _warned and optional:   # already known to be missing: do not walk sys.path again for every report
        return None
    try:
        import matplotlib

        if "matplotlib.pyplot" not in sys.modules:
            try:
                import os

                if not os.environ.get("DISPLAY") and not os.environ.get("MPLBACKEND"):
                    matplotlib.use("Agg")
            except Exception:
                pass
        import matplotlib.pyplot as plt

        return plt
    except ImportError:
        if not optional:
            raise
        if not _warned:
            print("note: matplotlib is not installed - figures are skipped, metrics are unaffected", file=sys.stderr)
            _warned = True
        return None


def create_figure_and_axis(title: Optional[str] = None, figure_size: Tuple[float, float] = DEFAULT_FIGURE_SIZE):
    plt = pyplot(optional=False)
    figure, axis = plt.subplots(figsize=figure_size, dpi=DEFAULT_DPI)
    if title is not None:
        axis.set_title(title)
    axis.grid(DEFAULT_GRID)
    return figure, axis


def finalize_and_show_or_save(figure, output_path: Optional[str | Path] = None, show_interactive: bool = True) -> None:
    """Save as PNG when a path is given, else show (unless disabled); the figure is closed either way."""
    plt = pyplot(optional=False)
    if output_path is not None:
        target = Path(output_path)
        target.parent.mkdir(parents=True, exist_ok=True)
        figure.savefig(target, bbox_inches="tight")
    elif show_interactive:
        plt.show()
    plt.close(figure)


def label_time_axis_seconds(axis) -> None:
    axis.set_xlabel("Time (seconds)")


def label_frequency_axis_hz(axis, log_scale: bool = False) -> None:
    axis.set_xlabel("Frequency (Hz)")
    if log_scale:
        axis.set_xscale("log")


def label_amplitude_axis(axis, unit: str = "Amplitude") -> None:
    axis.set_ylabel(unit)


def label_decibel_axis(axis) -> None:
    axis.set_ylabel("Level (dB)")
