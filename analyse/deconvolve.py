"""Drop-in module name: `analyse.deconvolve` is `audio_analysis_b200.deconvolve` (B200 deconvolution path)."""

import sys as _sys

from audio_analysis_b200 import deconvolve as _impl

_sys.modules[__name__] = _impl
