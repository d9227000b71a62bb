from .backbones.dit import DiT
from .cfm import CFM
from .duration_predictor import DurationPredictor
from .modules import MelSpec

__all__ = ["CFM", "DiT", "MelSpec", "DurationPredictor"]
