"""Drop-in module name: `analyse.modalcloud` is `audio_analysis_b200.modalcloud` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import modalcloud as _impl

_sys.modules[__name__] = _impl
