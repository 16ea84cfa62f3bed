"""String-valued option sets of the config surface.  The member names / values are the reference's
(dolomite_engine/hf_models/enums.py) because YAML files and saved `config.json`s carry them verbatim; they are built with
the functional Enum API from one table so that `Enum(value)` / `Enum[name]` / `.value` behave exactly like the reference's
classes."""

from enum import Enum

_OPTION_SETS = {
    # how weights are initialised (config.init_method)
    "InitMethod": ("normal", "mup"),
    # config.position_embedding_type; the B200 engine implements learned_absolute, rope (+ YaRN) and nope
    "PositionEmbeddingType": ("learned_absolute", "alibi", "rope", "nope"),
    # config.attention_head_type: layout of the fused c_attn output (attention/utils.py:18-106)
    "AttentionHeadType": ("mha", "mqa", "gqa"),
}


def _make(name: str) -> type[Enum]:
    return Enum(name, {v: v for v in _OPTION_SETS[name]}, module=__name__)


InitMethod = _make("InitMethod")
PositionEmbeddingType = _make("PositionEmbeddingType")
AttentionHeadType = _make("AttentionHeadType")
