"""Drop-in module name: `analyse.bundle` is `audio_analysis_b200.bundle` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import bundle as _impl

_sys.modules[__name__] = _impl
