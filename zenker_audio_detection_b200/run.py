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
    # The scripts threshold the returned scores on the host (ref:312-320, ref:333), so the drop-in models must know the
    # decision points to re-check around (model.py): pick the threshold flags out of the script's own command line.
    thr = []
    for flag in ("--stage1-threshold", "--stage2-threshold", "--stage1-forward-min-prob"):
        for i, a in enumerate(argv):
            v = argv[i + 1] if a == flag and i + 1 < len(argv) else (a.split("=", 1)[1] if a.startswith(flag + "=") else None)
            if v is not None:
                try:
                    thr.append(str(float(v)))
                except ValueError:
                    pass
    if thr and "ZK_RECHECK_THRESHOLDS" not in os.environ:
        os.environ["ZK_RECHECK_THRESHOLDS"] = ",".join(["0.5"] + thr)
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
