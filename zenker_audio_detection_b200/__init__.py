"""zenker-b200: the two-stage sliding-window inference path of daostler-tum/zenker-audio-detection on B200.

Public surface (drop-in for the reference's two call contracts):
    ZenkerASTFeatureExtractor        ~ transformers.ASTFeatureExtractor
    ZenkerASTForAudioClassification  ~ transformers.ASTForAudioClassification
    TwoStagePipeline                 the fused cascade (resample -> fbank -> Stage 1 -> gate -> Stage 2)
    compat.patch_transformers()      run the reference scripts unmodified on the B200 path
All device work goes through libzk_b200.so (include/zk_b200.h); importing this package does not load it,
the first compute call does and raises ``ZkError`` if it is missing.
"""
from ._lib import ZkError  # noqa: F401

__all__ = ["ZkError", "ZenkerASTFeatureExtractor", "ZenkerASTForAudioClassification", "TwoStagePipeline"]


def __getattr__(name):
    if name == "ZenkerASTFeatureExtractor":
        from .fx import ZenkerASTFeatureExtractor

        return ZenkerASTFeatureExtractor
    if name == "ZenkerASTForAudioClassification":
        from .model import ZenkerASTForAudioClassification

        return ZenkerASTForAudioClassification
    if name == "TwoStagePipeline":
        from .pipeline import TwoStagePipeline

        return TwoStagePipeline
    raise AttributeError(name)
