"""Enums of the reference config surface (hf_models/enums.py)."""

from enum import Enum


class InitMethod(Enum):
    normal = "normal"
    mup = "mup"


class PositionEmbeddingType(Enum):
    learned_absolute = "learned_absolute"
    alibi = "alibi"
    rope = "rope"
    nope = "nope"


class AttentionHeadType(Enum):
    mha = "mha"
    mqa = "mqa"
    gqa = "gqa"
