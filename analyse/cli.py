"""Drop-in module name: `python -m analyse.cli ...` runs `audio_analysis_b200.cli`."""

from audio_analysis_b200.cli import *  # noqa: F401,F403
from audio_analysis_b200.cli import main

if __name__ == "__main__":
    main()
