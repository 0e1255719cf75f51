"""Drop-in module name: `analyse.report` is `audio_analysis_b200.report` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import report as _impl

_sys.modules[__name__] = _impl
