from .f5tts_wrapper import F5TTSWrapper
from .utils_infer import chunk_text, load_checkpoint, load_model, load_vocoder

__all__ = ["F5TTSWrapper", "chunk_text", "load_checkpoint", "load_model", "load_vocoder"]
