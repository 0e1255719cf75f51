"""CPU oracle for the IR decay / spectral-decay hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`audio_analysis_b200`,
`analyse`) may import this.  Allowed importers: `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs, and there only as the
checker or as the timed CPU baseline, never as the thing shipped.

Parity status: PINNED against outputs of the reference itself, generated in the
build container by importing `/root/reference/analyse/*` unmodified through a
matplotlib stub (`tests/golden/make_golden.py`) and committed as
`tests/golden/*.npz`.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so those fixtures plus the analytic known answers in
`tests/test_oracle_known_answers.py` are the pin.
"""

from .ref_port import (  # noqa: F401
    DecayParams,
    BandParams,
    CsdParams,
    schroeder_edc_db,
    crossing_time,
    line_fit,
    decay_metrics,
    band_table,
    band_mask,
    band_filter,
    band_metrics,
    stft_db,
    pick_slices,
    csd_slices,
    waterfall,
    SpectrogramParams,
    spectrogram,
    ModalParams,
    modal_log_bins,
    modal_aggregate,
    modal_fit,
    modal_cloud,
    deconvolve,
)
