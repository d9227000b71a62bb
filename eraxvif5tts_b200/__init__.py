"""eraxvif5tts_b200 — B200-native (sm_100a) F5-TTS hot path behind the reference's Python API.

The classes mirror hungkq-1724/EraXviF5TTS (`F5TTSWrapper`, `CFM`, `DiT`, `MelSpec`, a Vocos-compatible vocoder) with the same
signatures and `state_dict` keys; all device math runs in the hand-written CUDA library `lib/libf5b200.so`
(C ABI in include/f5b200.h).  There is no CPU / PyTorch fallback: importing works anywhere, running needs a B200 and the built
library.
"""
from . import _lib  # noqa: F401
from .model import CFM, DiT, MelSpec  # noqa: F401
from .vocoder import Vocos  # noqa: F401

__all__ = ["CFM", "DiT", "MelSpec", "Vocos"]
