"""fbkst_b200 -- B200-native (sm_100a) speech-translation encoder hot path.

Host-side mirror of FBK-fairseq-ST's ``examples/speech_recognition`` encoder interface
(``ConvolutionalTransformerEncoder``) over the C-ABI CUDA library ``libfbkst_b200.so``.
There is no CPU or PyTorch fallback: every op raises if the library or a B200 is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["ops", "encoder", "data", "pipeline", "sharding", "augment", "criterion", "cross_attention"]
