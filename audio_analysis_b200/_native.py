"""ctypes binding of libaudio_analysis_b200.so (include/audio_analysis_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (or
`make -C audio_analysis_b200/csrc`).  There is no CPU fallback: if the library
or a CUDA device is missing, every compute entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

LIB_NAME = "libaudio_analysis_b200.so"
LIB_PATH = Path(__file__).resolve().parent / "csrc" / LIB_NAME

AA_MAX_FITS = 3
AA_FIT_COLS = 8
AA_MAX_BANDS = 64
AA_MAX_SLICES = 64
AA_MAX_ROW_SAMPLES = 1 << 24

AA_ROW_OK = 0
AA_ROW_TOO_SHORT = 1

AA_BAND_LOWPASS, AA_BAND_BANDPASS, AA_BAND_HIGHPASS = 0, 1, 2
AA_DBREF_GLOBAL_MAX, AA_DBREF_SLICE_MAX, AA_DBREF_NONE = 0, 1, 2

EXPORTED_SYMBOLS = (
    "aa_abi_version",
    "aa_last_error",
    "aa_device_count",
    "aa_kernel_launch_count",
    "aa_profile_enable",
    "aa_profile_collect",
    "aa_profile_ms",
    "aa_profile_scopes",
    "aa_peak_index_batch",
    "aa_decay_batch",
    "aa_curve_fits_batch",
    "aa_box_smooth_db_batch",
    "aa_bands_plan_create",
    "aa_bands_plan_destroy",
    "aa_bands_plan_kind",
    "aa_bands_get_decimation",
    "aa_bands_set_decimation",
    "aa_bands_classify",
    "aa_bands_workspace_bytes",
    "aa_rt60bands_batch",
    "aa_rt60bands_batch_ex",
    "aa_sos_workspace_bytes",
    "aa_rt60bands_sos_batch",
    "aa_csd_batch",
    "aa_csd_plan_batch",
    "aa_csd_smooth_batch",
    "aa_stft_db_batch",
    "aa_modal_cloud_batch",
    "aa_pcm16_to_float_batch",
    "aa_deconv_plan_create",
    "aa_deconv_plan_destroy",
    "aa_deconv_workspace_bytes",
    "aa_deconvolve_batch",
)


class DecayParams(C.Structure):
    _fields_ = [
        ("sample_rate_hz", C.c_int32),
        ("trim_to_peak", C.c_int32),
        ("ignore_leading_samples", C.c_int32),
        ("smoothing_window", C.c_int32),
        ("edc_floor_db", C.c_double),
        ("edc_epsilon", C.c_double),
        ("fit_lower_limit_db", C.c_double),
        ("n_fits", C.c_int32),
        ("want_early10", C.c_int32),
        ("fit_hi_db", C.c_double * AA_MAX_FITS),
        ("fit_lo_db", C.c_double * AA_MAX_FITS),
    ]


class Band(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("reserved", C.c_int32),
        ("low_edge_hz", C.c_double),
        ("high_edge_hz", C.c_double),
    ]


class BandsOptions(C.Structure):
    """aa_bands_options: per-call decimation / certification settings of the band stage (zeros = defaults)."""

    _fields_ = [
        ("max_decimation", C.c_int32),
        ("refine", C.c_int32),
        ("cert_safety", C.c_double),
        ("cert_window", C.c_double),
        ("cert_tol", C.c_double),
        ("cert_min_span", C.c_int32),
        ("reserved", C.c_int32),
    ]


class BandsStats(C.Structure):
    """aa_bands_stats: (row, band) units each refinement level processed."""

    _fields_ = [
        ("n_levels", C.c_int32),
        ("level_limit", C.c_int32 * 4),
        ("level_units", C.c_int64 * 4),
        ("n_chunks", C.c_int32),
        ("rows_per_chunk", C.c_int32),
    ]


class CsdParams(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32),
        ("hop_length", C.c_int32),
        ("use_hann_window", C.c_int32),
        ("bin_lo", C.c_int32),
        ("n_bins", C.c_int32),
        ("db_reference", C.c_int32),
        ("floor_db", C.c_double),
        ("dynamic_range_db", C.c_double),
    ]


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"audio_analysis_b200 native call failed ({code}): {message}")
        self.code = code


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.aa_abi_version.restype = C.c_int
    lib.aa_abi_version.argtypes = []
    lib.aa_last_error.restype = C.c_char_p
    lib.aa_last_error.argtypes = []
    lib.aa_device_count.restype = C.c_int
    lib.aa_device_count.argtypes = []
    lib.aa_kernel_launch_count.restype = C.c_longlong
    lib.aa_kernel_launch_count.argtypes = []
    lib.aa_profile_enable.restype = None
    lib.aa_profile_enable.argtypes = [C.c_int]
    lib.aa_profile_collect.restype = C.c_int
    lib.aa_profile_collect.argtypes = []
    lib.aa_profile_ms.restype = C.c_double
    lib.aa_profile_ms.argtypes = [C.c_char_p]
    lib.aa_profile_scopes.restype = C.c_longlong
    lib.aa_profile_scopes.argtypes = [C.c_char_p]
    lib.aa_peak_index_batch.restype = C.c_int
    lib.aa_peak_index_batch.argtypes = [vp, i64, vp, i32, vp, vp]
    lib.aa_decay_batch.restype = C.c_int
    lib.aa_decay_batch.argtypes = [vp, i64, vp, i32, i32, C.POINTER(DecayParams), vp, vp, i64, vp, vp, vp, vp, vp]
    lib.aa_curve_fits_batch.restype = C.c_int
    lib.aa_curve_fits_batch.argtypes = [vp, i64, vp, i32, vp, i32, C.POINTER(DecayParams), vp, vp, vp]
    lib.aa_box_smooth_db_batch.restype = C.c_int
    lib.aa_box_smooth_db_batch.argtypes = [vp, vp, i64, vp, i32, vp, i32, i32, C.c_double, vp]
    lib.aa_bands_plan_create.restype = C.c_int
    lib.aa_bands_plan_create.argtypes = [i32, C.POINTER(vp)]
    lib.aa_bands_plan_destroy.restype = None
    lib.aa_bands_plan_destroy.argtypes = [vp]
    lib.aa_bands_plan_kind.restype = C.c_int
    lib.aa_bands_plan_kind.argtypes = [vp]
    lib.aa_bands_get_decimation.restype = C.c_int
    lib.aa_bands_get_decimation.argtypes = []
    lib.aa_bands_set_decimation.restype = None
    lib.aa_bands_set_decimation.argtypes = [C.c_int]
    lib.aa_bands_classify.restype = C.c_int
    lib.aa_bands_classify.argtypes = [vp, C.POINTER(Band), i32, C.c_double, i32, C.POINTER(DecayParams), C.POINTER(C.c_int32)]
    lib.aa_bands_workspace_bytes.restype = sz
    lib.aa_bands_workspace_bytes.argtypes = [vp, i32, i32]
    lib.aa_rt60bands_batch.restype = C.c_int
    lib.aa_rt60bands_batch.argtypes = [vp, vp, i64, i32, vp, C.POINTER(Band), i32, C.c_double,
                                       C.POINTER(DecayParams), vp, vp, vp, sz, vp]
    lib.aa_rt60bands_batch_ex.restype = C.c_int
    lib.aa_rt60bands_batch_ex.argtypes = [vp, vp, i64, i32, vp, C.POINTER(Band), i32, C.c_double,
                                          C.POINTER(DecayParams), vp, vp, vp, sz, vp, C.POINTER(BandsOptions),
                                          C.POINTER(BandsStats)]
    lib.aa_sos_workspace_bytes.restype = sz
    lib.aa_sos_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.aa_rt60bands_sos_batch.restype = C.c_int
    lib.aa_rt60bands_sos_batch.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, i32, i32, C.POINTER(DecayParams), vp, vp,
                                           vp, sz, vp]
    lib.aa_csd_batch.restype = C.c_int
    lib.aa_csd_batch.argtypes = [vp, i64, vp, i32, i32, vp, vp, vp, i32, C.POINTER(CsdParams), vp, vp]
    lib.aa_csd_plan_batch.restype = C.c_int
    lib.aa_csd_plan_batch.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.aa_csd_smooth_batch.restype = C.c_int
    lib.aa_csd_smooth_batch.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, i32, C.c_double, vp]
    lib.aa_stft_db_batch.restype = C.c_int
    lib.aa_stft_db_batch.argtypes = [vp, i64, vp, i32, i32, vp, vp, i32, i32, i32, C.c_double, i32, vp, vp]
    lib.aa_modal_cloud_batch.restype = C.c_int
    lib.aa_modal_cloud_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, i32, C.c_double, C.c_double, C.c_double,
                                         i32, C.c_double, C.c_double, vp, vp, vp]
    lib.aa_pcm16_to_float_batch.restype = C.c_int
    lib.aa_pcm16_to_float_batch.argtypes = [vp, i64, i32, i32, i32, i32, vp, i64, vp]
    lib.aa_deconv_plan_create.restype = C.c_int
    lib.aa_deconv_plan_create.argtypes = [i32, C.POINTER(vp)]
    lib.aa_deconv_plan_destroy.restype = None
    lib.aa_deconv_plan_destroy.argtypes = [vp]
    lib.aa_deconv_workspace_bytes.restype = sz
    lib.aa_deconv_workspace_bytes.argtypes = [vp, i32]
    lib.aa_deconvolve_batch.restype = C.c_int
    lib.aa_deconvolve_batch.argtypes = [vp, vp, i32, vp, i64, i32, i32, C.c_double, i32, C.c_double, i32, i32, vp,
                                        i64, vp, sz, vp]


def library_path() -> Path:
    override = os.environ.get("AUDIO_ANALYSIS_B200_LIB")
    return Path(override) if override else LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {LIB_PATH.parent}`. There is no CPU fallback."
            )
        lib = C.CDLL(str(path))
        _declare(lib)
        if lib.aa_abi_version() != 1:
            raise RuntimeError(f"{path}: unexpected ABI version {lib.aa_abi_version()}")
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = load().aa_last_error()
        raise NativeError(code, msg.decode("utf-8", "replace") if msg else "")


def require_device() -> None:
    n = load().aa_device_count()
    if n <= 0:
        raise RuntimeError(
            "audio_analysis_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback"
        )
