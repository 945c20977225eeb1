"""CPU oracle for the two-stage sliding-window inference path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from this package.  The product path
(``zenker_audio_detection_b200``) never imports it and fails loudly when the CUDA
library is missing.

Layout
------
``glue.py``        numpy restatement of the reference-owned glue
                   (``src/test_long_audio_windows_2stage.py`` and ``..._cache.py``):
                   window slicing, the Stage-1 gate, Stage-2 class vector,
                   ``summarize_stage_outputs`` and the patient aggregate.
``numerics.py``    numpy / torch-fp32 restatement of the third-party numerics the
                   reference calls (torchaudio ``resample`` + ``kaldi.fbank``,
                   HF ``ASTFeatureExtractor`` and ``ASTForAudioClassification``).
``thirdparty.py``  thin wrappers that call the *installed* torchaudio / transformers
                   (the packages the reference itself imports; they are part of the
                   image on the GPU box, ``/root/reference`` is not).

Parity pinning
--------------
The reference ships no tests and no golden vectors, so parity is
pinned by (a) fixtures under ``tests/golden/`` generated HERE by importing the
reference's own functions from ``/root/reference/src`` (script:
``scripts/make_golden.py``) and (b) the installed third-party packages executed on
CPU in the same test run.  ``tests/test_oracle_*.py`` checks the restatement against
both.
"""
