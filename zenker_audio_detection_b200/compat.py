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

    if "ASTFeatureExtractor" not in _originals:
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
        for k in ("ASTFeatureExtractor", "ASTForAudioClassification"):
            if k in _originals:
                setattr(m, k, _originals[k])
    for k in ("ASTFeatureExtractor", "ASTForAudioClassification"):
        _originals.pop(k, None)


def patch_torchaudio() -> None:
    """``torchaudio.load`` / ``torchaudio.info`` need TorchCodec in torchaudio >= 2.9 (``info`` is gone altogether) and
    fail in this image; the reference calls them in ``load_audio`` (ref:54) and ``discover_two_files`` (ref:132).
    Route both through the RIFF reader for ``.wav`` files: same return contract ``(float32 (channels, frames), sr)`` /
    an object with ``num_frames``."""
    import torch
    import torchaudio

    from . import wavio

    if "torchaudio.load" not in _originals:
        _originals["torchaudio.load"] = getattr(torchaudio, "load", None)
        _originals["torchaudio.info"] = getattr(torchaudio, "info", None)

    def load(path, *args, **kwargs):
        data, sr = wavio.read(str(path))
        if data.dtype.kind == "i":  # PCM16 (frames, channels) -> torchaudio's normalised float32 (channels, frames)
            data = (data.astype("float32") / 32768.0).T
        return torch.from_numpy(data.copy() if not data.flags.writeable else data).contiguous(), sr

    def info(path, *args, **kwargs):
        return wavio.info(str(path))

    torchaudio.load = load
    torchaudio.info = info


def unpatch_torchaudio() -> None:
    import torchaudio

    for k in ("load", "info"):
        v = _originals.pop(f"torchaudio.{k}", None)
        if v is not None:
            setattr(torchaudio, k, v)
        elif hasattr(torchaudio, k) and k == "info":
            delattr(torchaudio, k)
