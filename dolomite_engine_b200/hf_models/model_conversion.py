"""`import_from_huggingface(name_or_path, save_path)` for llama / granite checkpoints (reference:
hf_models/model_conversion/__init__.py:19-27, llama.py:13-149, granite.py:15-80).  CPU-only weight re-layout:

  * q/k/v projections -> one `c_attn` in the per-head / per-group interleave of attention/utils.py:18-106
  * up_proj / gate_proj -> `c_fc = cat([up, gate])` (gpt_dolomite/mlp.py:54-55)
  * LlamaConfig / GraniteConfig -> GPTDolomiteConfig (llama.py:37-74, granite.py:40-79)

Works on local directories (config.json + *.safetensors); there is no hub access in this environment.
Export and the other model families are 'next' rows.
"""

from __future__ import annotations

import json
import os

import torch

from ..utils.safetensors import SafeTensorsWeightsManager
from .config import GPTDolomiteConfig


def interleave_query_key_value_tensor_for_attention(q, k, v, num_heads: int, num_key_value_heads: int, head_dim: int,
                                                    attention_head_type: str) -> torch.Tensor:
    """attention/__init__.py:62-84 -> utils.py:18-106.  Accepts tensors or safetensors slices."""
    q, k, v = q[:], k[:], v[:]
    if attention_head_type == "mqa":
        return torch.cat([q, k, v])
    g = num_heads // num_key_value_heads
    parts = []
    for i in range(num_key_value_heads):
        parts.append(q[i * g * head_dim : (i + 1) * g * head_dim])
        parts.append(k[i * head_dim : (i + 1) * head_dim])
        parts.append(v[i * head_dim : (i + 1) * head_dim])
    return torch.cat(parts)


def split_query_key_value_tensor_for_attention(w, num_heads: int, num_key_value_heads: int, head_dim: int,
                                               attention_head_type: str):
    if attention_head_type == "mqa":
        return w.split((num_heads * head_dim, head_dim, head_dim))
    g = num_heads // num_key_value_heads
    x = w.view(num_key_value_heads, g + 2, head_dim, *w.shape[1:])
    return (x[:, :g].reshape(-1, *w.shape[1:]), x[:, g].reshape(-1, *w.shape[1:]), x[:, g + 1].reshape(-1, *w.shape[1:]))


def interleave_up_gate_tensor_for_mlp(up: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    return torch.cat([up, gate])


def split_up_gate_tensor_for_mlp(c_fc: torch.Tensor):
    return c_fc.chunk(2)


def _head_type(cfg: dict) -> str:
    nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads") or cfg["num_attention_heads"]
    if nh == nkv:
        return "mha"
    if nkv == 1:
        return "mqa"
    assert nh > nkv
    return "gqa"


def _import_config(original: dict) -> GPTDolomiteConfig:
    assert original.get("hidden_act", "silu") == "silu"
    assert original.get("mlp_bias", False) == original.get("attention_bias", False)
    kw = dict(
        vocab_size=original["vocab_size"],
        n_positions=original["max_position_embeddings"],
        n_embd=original["hidden_size"],
        n_layer=original["num_hidden_layers"],
        n_head=original["num_attention_heads"],
        num_key_value_heads=original.get("num_key_value_heads") or original["num_attention_heads"],
        attention_head_type=_head_type(original),
        position_embedding_type="rope",
        n_inner=original["intermediate_size"],
        activation_function="swiglu",
        normalization_function="rmsnorm",
        layer_norm_epsilon=original.get("rms_norm_eps", 1e-6),
        use_cache=original.get("use_cache", True),
        add_bias=original.get("attention_bias", False),
        tie_word_embeddings=original.get("tie_word_embeddings", False),
        initializer_range=original.get("initializer_range", 0.02),
        rope_theta=original.get("rope_theta", 10000),
        rope_scaling=original.get("rope_scaling"),
        attn_pdrop=original.get("attention_dropout", 0.0),
        resid_pdrop=0.0,
        embd_pdrop=0.0,
        bos_token_id=original.get("bos_token_id"),
        eos_token_id=original.get("eos_token_id"),
        pad_token_id=original.get("pad_token_id"),
    )
    if original["model_type"] == "granite":
        one = lambda x: None if x == 1 else x  # noqa: E731
        kw.update(
            m_emb=one(original.get("embedding_multiplier", 1)),
            m_residual=one(original.get("residual_multiplier", 1)),
            m_width=one(original.get("logits_scaling", 1)),
            attention_multiplier=original.get("attention_multiplier"),
        )
    return GPTDolomiteConfig(**kw)


def _import_state_dict(m: SafeTensorsWeightsManager, config: GPTDolomiteConfig) -> dict:
    hd = config.n_embd // config.n_head
    sd = {
        "transformer.wte.weight": m.get_tensor("model.embed_tokens.weight"),
        "transformer.ln_f.weight": m.get_tensor("model.norm.weight"),
    }
    if m.has_tensor("lm_head.weight") and not config.tie_word_embeddings:
        sd["lm_head.weight"] = m.get_tensor("lm_head.weight")
    for i in range(config.n_layer):
        src, dst = f"model.layers.{i}.", f"transformer.h.{i}."
        sd[dst + "ln_1.weight"] = m.get_tensor(src + "input_layernorm.weight")
        sd[dst + "ln_2.weight"] = m.get_tensor(src + "post_attention_layernorm.weight")
        for suffix in ("weight", "bias"):
            if not m.has_tensor(src + f"mlp.up_proj.{suffix}"):
                continue
            sd[dst + f"mlp.c_fc.{suffix}"] = interleave_up_gate_tensor_for_mlp(
                m.get_tensor(src + f"mlp.up_proj.{suffix}"), m.get_tensor(src + f"mlp.gate_proj.{suffix}"))
            sd[dst + f"mlp.c_proj.{suffix}"] = m.get_tensor(src + f"mlp.down_proj.{suffix}")
            sd[dst + f"attn.c_attn.{suffix}"] = interleave_query_key_value_tensor_for_attention(
                m.get_tensor(src + f"self_attn.q_proj.{suffix}"), m.get_tensor(src + f"self_attn.k_proj.{suffix}"),
                m.get_tensor(src + f"self_attn.v_proj.{suffix}"), config.n_head, config.num_key_value_heads, hd,
                config.attention_head_type)
            sd[dst + f"attn.c_proj.{suffix}"] = m.get_tensor(src + f"self_attn.o_proj.{suffix}")
    return sd


_SUPPORTED = ("llama", "granite")


def import_from_huggingface(pretrained_model_name_or_path: str, save_path: str) -> None:
    if not os.path.isdir(pretrained_model_name_or_path):
        raise FileNotFoundError(
            f"{pretrained_model_name_or_path}: only local checkpoints can be imported (no hub access in this environment)")
    with open(os.path.join(pretrained_model_name_or_path, "config.json")) as f:
        original = json.load(f)
    model_type = original.get("model_type")
    if model_type not in _SUPPORTED:
        raise NotImplementedError(f"the current model_type ({model_type}) is not yet supported")
    config = _import_config(original)
    sd = _import_state_dict(SafeTensorsWeightsManager(pretrained_model_name_or_path), config)
    SafeTensorsWeightsManager.save_state_dict(sd, save_path)
    config.save_pretrained(save_path)
    for extra in ("tokenizer.json", "tokenizer_config.json", "special_tokens_map.json", "tokenizer.model"):
        p = os.path.join(pretrained_model_name_or_path, extra)
        if os.path.exists(p):
            import shutil

            shutil.copy(p, os.path.join(save_path, extra))
