"""Drop-in surface of `dolomite_engine.hf_models` for the data-parallel training hot path."""

from .config import CommonConfig, GPTDolomiteConfig, MoEDolomiteConfig, config_class_for, config_for_model
from .enums import AttentionHeadType, InitMethod, PositionEmbeddingType
from .modeling import (
    AutoModelForCausalLM,
    CausalLMOutputWithPast,
    DolomitePreTrainedModel,
    GPTDolomiteForCausalLM,
    MoEDolomiteForCausalLM,
)
from .utils import convert_padding_free_lists_to_tensors
from .model_conversion import export_to_huggingface, import_from_huggingface
