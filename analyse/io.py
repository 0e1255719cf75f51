"""Drop-in module name: `analyse.io` is `audio_analysis_b200.io` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import io as _impl

_sys.modules[__name__] = _impl
