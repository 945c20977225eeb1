"""``python -m zenker_audio_detection_b200.run <reference script.py> [args...]`` -- run an unmodified reference
script (e.g. src/test_long_audio_windows_2stage.py) with its feature extractor and model replaced by the B200 path."""
from __future__ import annotations

import os
import runpy
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m zenker_audio_detection_b200.run <script.py> [script args...]")
    from .compat import patch_torchaudio, patch_transformers

    patch_transformers()
    patch_torchaudio()  # WAV decoding without TorchCodec (ref:54, ref:132)
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
