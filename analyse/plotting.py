"""Drop-in module name: `analyse.plotting` is `audio_analysis_b200.plotting` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import plotting as _impl

_sys.modules[__name__] = _impl
