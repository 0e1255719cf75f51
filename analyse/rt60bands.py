"""Drop-in module name: `analyse.rt60bands` is `audio_analysis_b200.rt60bands` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import rt60bands as _impl

_sys.modules[__name__] = _impl
