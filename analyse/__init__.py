"""`analyse` - the reference's package name, served by the B200 implementation.

Only the accelerated path is provided: `analyse.io`, `analyse.decay`, `analyse.rt60bands`,
`analyse.waterfall`, `analyse.spectrogram`, `analyse.modalcloud`,
`analyse.bundle`, `analyse.report`, `analyse.cli`, `analyse.plotting`.
"""

from audio_analysis_b200 import *  # noqa: F401,F403
from audio_analysis_b200 import __all__  # noqa: F401
