"""Drop-in module name: `analyse.decay` is `audio_analysis_b200.decay` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import decay as _impl

_sys.modules[__name__] = _impl
