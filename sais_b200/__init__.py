"""sais_b200 — B200-native (sm_100a) implementation of SAIS's data-parallel inference hot path:
DINO ViT-S/16 per-frame features -> temporal TransformerEncoder returning (output, attn) -> prototype scoring.

Module names mirror the reference scripts so that ``import sais_b200.vision_transformer as vits`` and
``from sais_b200.prepare_model import loadModel`` replace the reference imports one for one.
"""
from . import _lib  # noqa: F401
from ._lib import SaisError, build, lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["SaisError", "build", "lib", "vision_transformer", "prepare_model", "transformer", "scoring", "ops",
           "pipeline"]
