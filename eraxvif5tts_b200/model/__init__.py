from .backbones.dit import DiT
from .cfm import CFM
from .modules import MelSpec

__all__ = ["CFM", "DiT", "MelSpec"]
