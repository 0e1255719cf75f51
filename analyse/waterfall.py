"""Drop-in module name: `analyse.waterfall` is `audio_analysis_b200.waterfall` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import waterfall as _impl

_sys.modules[__name__] = _impl
