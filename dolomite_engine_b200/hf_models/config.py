"""Config surface of the reference (hf_models/config.py:6-111, models/gpt_dolomite/config.py:4-5,
models/moe_dolomite/config.py:4-83): same field names, defaults, aliases, validation and config.json format.

The reference subclasses transformers.PretrainedConfig; that class changed incompatibly in transformers 5.x
(SURVEY.md section 8c), so this is a small self-contained implementation of the same surface:
`to_dict / from_dict / save_pretrained / from_pretrained`, GPT-2 style names plus the `attribute_map` aliases.
"""

from __future__ import annotations

import copy
import json
import os

from .enums import AttentionHeadType, InitMethod, PositionEmbeddingType


class CommonConfig:
    model_type = "common"
    keys_to_ignore_at_inference = ["past_key_values"]
    attribute_map = {
        "hidden_size": "n_embd",
        "max_position_embeddings": "n_positions",
        "num_attention_heads": "n_head",
        "num_hidden_layers": "n_layer",
    }

    def __init__(
        self,
        vocab_size: int = 50257,
        n_positions: int = 1024,
        n_embd: int = 768,
        n_layer: int = 12,
        n_head: int = 12,
        num_key_value_heads: int | None = None,
        n_inner: int | None = None,
        activation_function: str = "gelu_pytorch_tanh",
        attention_head_type: str = "mqa",
        resid_pdrop: float = 0.1,
        embd_pdrop: float = 0.1,
        attn_pdrop: float = 0.1,
        normalization_function: str = "layernorm",
        layer_norm_epsilon: float = 1e-5,
        initializer_range: float = 0.02,
        scale_attn_weights: bool = True,
        attention_multiplier: float | None = None,
        use_cache: bool = True,
        bos_token_id: int = 50256,
        eos_token_id: int = 50256,
        pad_token_id: int = 50256,
        attention_softmax_in_fp32: bool = True,
        add_bias: bool = True,
        position_embedding_type: str = "learned_absolute",
        rope_theta: int = 10000,
        rope_scaling: dict | None = None,
        m_emb: float | None = None,
        m_width: float | None = None,
        m_residual: float | None = None,
        init_method: str = "normal",
        upcast_logits_for_loss: bool = False,
        tie_word_embeddings: bool = True,
        **kwargs,
    ) -> None:
        # aliases passed as kwargs (hidden_size=..., num_hidden_layers=...) map onto the GPT-2 names
        for alias, name in self.attribute_map.items():
            if alias in kwargs:
                val = kwargs.pop(alias)
                if name == "n_embd":
                    n_embd = val
                elif name == "n_positions":
                    n_positions = val
                elif name == "n_head":
                    n_head = val
                elif name == "n_layer":
                    n_layer = val
        self.vocab_size = vocab_size
        self.n_positions = n_positions
        self.n_embd = n_embd
        self.n_layer = n_layer
        self.n_head = n_head
        self.num_key_value_heads = num_key_value_heads
        self.n_inner = 4 * n_embd if n_inner is None else n_inner
        self.activation_function = activation_function
        self.attention_head_type = attention_head_type
        self.resid_pdrop = resid_pdrop
        self.embd_pdrop = embd_pdrop
        self.attn_pdrop = attn_pdrop
        self.normalization_function = normalization_function
        self.layer_norm_epsilon = layer_norm_epsilon
        self.initializer_range = initializer_range
        self.scale_attn_weights = scale_attn_weights
        self.attention_multiplier = attention_multiplier
        self.use_cache = use_cache
        self.attention_softmax_in_fp32 = attention_softmax_in_fp32
        self.position_embedding_type = position_embedding_type
        self.add_bias = add_bias
        self.rope_theta = rope_theta
        self.rope_scaling = rope_scaling
        self.m_emb = m_emb
        self.m_width = m_width
        self.m_residual = m_residual
        self.init_method = init_method
        self.upcast_logits_for_loss = upcast_logits_for_loss
        self.tie_word_embeddings = tie_word_embeddings
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.pad_token_id = pad_token_id

        if self.attention_multiplier is not None:
            assert self.scale_attn_weights

        # check if enums are valid
        init_method = InitMethod(init_method)
        attention_head_type = AttentionHeadType(attention_head_type)
        position_embedding_type = PositionEmbeddingType(position_embedding_type)

        self.multi_query = attention_head_type == AttentionHeadType.mqa

        if attention_head_type == AttentionHeadType.mha:
            if self.num_key_value_heads is None:
                self.num_key_value_heads = self.n_head
            assert (
                self.n_head == self.num_key_value_heads
            ), "MultiHeadAttention should have same number of heads for query, keys and values"
        elif attention_head_type == AttentionHeadType.mqa:
            if self.num_key_value_heads is None:
                self.num_key_value_heads = 1
            assert self.num_key_value_heads == 1, "MultiQueryAttention should have 1 head for keys and values"
        elif attention_head_type == AttentionHeadType.gqa:
            assert (
                self.num_key_value_heads is not None
            ), "`num_key_value_heads` needs to be specified with GroupedQueryAttention"
            assert (
                self.n_head % self.num_key_value_heads == 0
            ), "GroupedQueryAttention should have more than 1 head for keys and values"

        self._extra = dict(kwargs)  # unknown HF keys are kept so that config.json round-trips

    # ---- attribute_map aliases ----
    def __getattr__(self, name):
        amap = type(self).attribute_map
        if name in amap:
            return getattr(self, amap[name])
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    def __setattr__(self, name, value):
        amap = type(self).attribute_map
        object.__setattr__(self, amap.get(name, name), value)

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
        d = dict(d)
        d.pop("model_type", None)
        d.pop("multi_query", None)
        for k in ("architectures", "transformers_version", "torch_dtype", "dtype"):
            d.pop(k, None)
        return cls(**d)

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

    def __init__(
        self,
        num_experts: int = 8,
        num_experts_per_tok: int = 2,
        output_router_logits: bool = False,
        router_aux_loss_coef: float = 0.001,
        **kwargs,
    ) -> None:
        super().__init__(**kwargs)
        self.num_experts = num_experts
        self.num_experts_per_tok = num_experts_per_tok
        self.output_router_logits = output_router_logits
        self.router_aux_loss_coef = router_aux_loss_coef


_CONFIG_CLASSES = {"gpt_dolomite": GPTDolomiteConfig, "moe_dolomite": MoEDolomiteConfig}


def config_class_for(model_type: str):
    if model_type not in _CONFIG_CLASSES:
        raise ValueError(f"unexpected model_type ({model_type}); the B200 path implements {sorted(_CONFIG_CLASSES)}")
    return _CONFIG_CLASSES[model_type]


def config_for_model(model_type: str, **kwargs):
    """AutoConfig.for_model(**pretrained_config) equivalent (model_wrapper/base.py:151-163)"""
    return config_class_for(model_type)(**kwargs)
