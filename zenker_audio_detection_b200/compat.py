"""Swap mechanism: the reference scripts do
``from transformers import ASTFeatureExtractor, ASTConfig, ASTForAudioClassification`` (ref:40); rebinding the two
class names on the ``transformers`` module before the script is imported routes its hot path through libzk_b200
while everything else (``ASTConfig``, argparse, JSON writing) stays the reference's own code."""
from __future__ import annotations

_originals = {}


def _modules():
    """``import transformers`` may hand back the bootstrap module while ``sys.modules`` holds the lazy proxy that
    ``from transformers import X`` consults -- patch every distinct object."""
    import sys

    import transformers

    mods = [transformers]
    if sys.modules.get("transformers") is not transformers:
        mods.append(sys.modules["transformers"])
    return mods


def patch_transformers() -> None:
    from .fx import ZenkerASTFeatureExtractor
    from .model import ZenkerASTForAudioClassification

    if not _originals:
        # resolving the lazy attributes can REPLACE sys.modules["transformers"] (observed with 5.5.0), so
        # resolve first, collect the module objects afterwards
        import transformers

        _originals["ASTFeatureExtractor"] = transformers.ASTFeatureExtractor
        _originals["ASTForAudioClassification"] = transformers.ASTForAudioClassification
    mods = _modules()
    for m in mods:
        m.ASTFeatureExtractor = ZenkerASTFeatureExtractor
        m.ASTForAudioClassification = ZenkerASTForAudioClassification


def unpatch_transformers() -> None:
    for m in _modules():
        for k, v in _originals.items():
            setattr(m, k, v)
    _originals.clear()
