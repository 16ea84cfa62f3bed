"""Config surface of the reference (hf_models/config.py:6-111, models/gpt_dolomite/config.py:4-5,
models/moe_dolomite/config.py:4-83): same field names, defaults, aliases, validation and config.json format.

The reference subclasses transformers.PretrainedConfig; that class changed incompatibly in transformers 5.x
(SURVEY.md section 8c), so this is a small self-contained, table-driven implementation of the same surface:
every config class lists its (field, default) pairs once in `FIELDS`; construction, `to_dict / from_dict /
save_pretrained / from_pretrained` and the `attribute_map` aliases are generic.
"""

from __future__ import annotations

import copy
import json
import os

from .enums import AttentionHeadType, InitMethod, PositionEmbeddingType

# HuggingFace-style aliases of the GPT-2 style names (config.py:8-13)
_ALIASES = {"hidden_size": "n_embd", "max_position_embeddings": "n_positions", "num_attention_heads": "n_head",
            "num_hidden_layers": "n_layer"}
# keys HuggingFace writes into config.json that are not model hyper-parameters
_HF_NOISE = ("model_type", "multi_query", "architectures", "transformers_version", "torch_dtype", "dtype")


class CommonConfig:
    model_type = "common"
    keys_to_ignore_at_inference = ["past_key_values"]
    attribute_map = _ALIASES

    # (name, default) in the reference's order; None defaults are resolved in `_finalise`
    FIELDS: tuple[tuple[str, object], ...] = (
        ("vocab_size", 50257), ("n_positions", 1024), ("n_embd", 768), ("n_layer", 12), ("n_head", 12),
        ("num_key_value_heads", None), ("n_inner", None), ("activation_function", "gelu_pytorch_tanh"),
        ("attention_head_type", "mqa"), ("resid_pdrop", 0.1), ("embd_pdrop", 0.1), ("attn_pdrop", 0.1),
        ("normalization_function", "layernorm"), ("layer_norm_epsilon", 1e-5), ("initializer_range", 0.02),
        ("scale_attn_weights", True), ("attention_multiplier", None), ("use_cache", True), ("bos_token_id", 50256),
        ("eos_token_id", 50256), ("pad_token_id", 50256), ("attention_softmax_in_fp32", True), ("add_bias", True),
        ("position_embedding_type", "learned_absolute"), ("rope_theta", 10000), ("rope_scaling", None), ("m_emb", None),
        ("m_width", None), ("m_residual", None), ("init_method", "normal"), ("upcast_logits_for_loss", False),
        ("tie_word_embeddings", True),
    )

    def __init__(self, **kwargs) -> None:
        known = dict(type(self).all_fields())
        values = {}
        for key in list(kwargs):
            name = _ALIASES.get(key, key)
            if name in known:
                values[name] = kwargs.pop(key)
        for name, default in known.items():
            object.__setattr__(self, name, values.get(name, copy.deepcopy(default)))
        self._extra = dict(kwargs)  # unknown HF keys are kept so that config.json round-trips
        self._finalise()

    @classmethod
    def all_fields(cls) -> tuple[tuple[str, object], ...]:
        out: tuple[tuple[str, object], ...] = ()
        for klass in reversed(cls.__mro__):
            out += tuple(klass.__dict__.get("FIELDS", ()))
        return out

    def _finalise(self) -> None:
        """derived defaults and the reference's consistency checks (config.py:56, :81-109)"""
        if self.n_inner is None:
            self.n_inner = 4 * self.n_embd
        if self.attention_multiplier is not None and not self.scale_attn_weights:
            raise AssertionError("attention_multiplier needs scale_attn_weights")
        InitMethod(self.init_method)  # ValueError for unknown members, like the reference's enum casts
        PositionEmbeddingType(self.position_embedding_type)
        kind = AttentionHeadType(self.attention_head_type)
        self.multi_query = kind is AttentionHeadType.mqa
        nkv = self.num_key_value_heads
        if kind is AttentionHeadType.mha:
            nkv = self.n_head if nkv is None else nkv
            assert nkv == self.n_head, "MultiHeadAttention should have same number of heads for query, keys and values"
        elif kind is AttentionHeadType.mqa:
            nkv = 1 if nkv is None else nkv
            assert nkv == 1, "MultiQueryAttention should have 1 head for keys and values"
        else:
            assert nkv is not None, "`num_key_value_heads` needs to be specified with GroupedQueryAttention"
            assert self.n_head % nkv == 0, "GroupedQueryAttention should have more than 1 head for keys and values"
        self.num_key_value_heads = nkv

    # ---- attribute_map aliases ----
    def __getattr__(self, name):
        if name in _ALIASES:
            return getattr(self, _ALIASES[name])
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    def __setattr__(self, name, value):
        object.__setattr__(self, _ALIASES.get(name, name), value)

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head

    # ---- (de)serialisation ----
    def to_dict(self) -> dict:
        d = {k: copy.deepcopy(v) for k, v in self.__dict__.items() if not k.startswith("_")}
        d.update(copy.deepcopy(self._extra))
        d["model_type"] = self.model_type
        return d

    @classmethod
    def from_dict(cls, d: dict):
        return cls(**{k: v for k, v in d.items() if k not in _HF_NOISE})

    def save_pretrained(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)

    @classmethod
    def from_pretrained(cls, path: str):
        with open(os.path.join(path, "config.json")) as f:
            d = json.load(f)
        return config_class_for(d.get("model_type", cls.model_type)).from_dict(d)

    def __repr__(self) -> str:
        return f"{type(self).__name__} {json.dumps(self.to_dict(), indent=2, sort_keys=True)}"


class GPTDolomiteConfig(CommonConfig):
    model_type = "gpt_dolomite"


class MoEDolomiteConfig(CommonConfig):
    """models/moe_dolomite/config.py:4-83"""

    model_type = "moe_dolomite"
    FIELDS = (("num_experts", 8), ("num_experts_per_tok", 2), ("output_router_logits", False), ("router_aux_loss_coef", 0.001))

    def _finalise(self) -> None:
        super()._finalise()
        assert self.init_method == "normal", "MoEDolomite supports the normal init method only (moe_dolomite/config.py:83)"


_CONFIG_CLASSES = {c.model_type: c for c in (GPTDolomiteConfig, MoEDolomiteConfig)}


def config_class_for(model_type: str):
    if model_type not in _CONFIG_CLASSES:
        raise ValueError(f"unexpected model_type ({model_type}); the B200 path implements {sorted(_CONFIG_CLASSES)}")
    return _CONFIG_CLASSES[model_type]


def config_for_model(model_type: str, **kwargs):
    """AutoConfig.for_model(**pretrained_config) equivalent (model_wrapper/base.py:151-163)"""
    return config_class_for(model_type)(**kwargs)
