from .backbones.dit import DiT
from .cfm import CFM
from .duration_predictor import DurationPredictor
from .modules import MelSpec
from .trainer import Trainer

__all__ = ["CFM", "DiT", "MelSpec", "DurationPredictor", "Trainer"]
