"""Drop-in module name: `analyse.spectrogram` is `audio_analysis_b200.spectrogram` (B200 decay path)."""

import sys as _sys

from audio_analysis_b200 import spectrogram as _impl

_sys.modules[__name__] = _impl
