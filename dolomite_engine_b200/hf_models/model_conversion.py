"""`import_from_huggingface(name_or_path, save_path)` / `export_to_huggingface` (reference:
hf_models/model_conversion/__init__.py:11-46).  llama / granite live here (llama.py:13-290, granite.py:15-142); gpt_bigcode,
mixtral and granitemoe in model_conversion_families.py.  CPU-only weight re-layout:

  * q/k/v projections -> one `c_attn` in the per-head / per-group interleave of attention/utils.py:18-106
  * up_proj / gate_proj -> `c_fc = cat([up, gate])` (gpt_dolomite/mlp.py:54-55)
  * LlamaConfig / GraniteConfig -> GPTDolomiteConfig (llama.py:37-74, granite.py:40-79)

Works on local directories (config.json + *.safetensors); there is no hub access in this environment.
"""

from __future__ import annotations

import json
import os

import torch

from ..utils.safetensors import SafeTensorsWeightsManager
from . import model_conversion_families as F
from .config import CommonConfig, GPTDolomiteConfig


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


def rope_fields(original: dict) -> tuple[float, dict | None]:
    """(rope_theta, rope_scaling) from either config dialect: the top-level `rope_theta` / `rope_scaling` keys the reference
    reads (llama.py:60-61), or the `rope_parameters` dict transformers >= 5 writes ({"rope_theta", "rope_type", ...})"""
    rp = original.get("rope_parameters")
    if rp is None:
        return original.get("rope_theta", 10000), original.get("rope_scaling")
    theta = rp.get("rope_theta", original.get("rope_theta", 10000))
    kind = rp.get("rope_type", rp.get("type", "default"))
    if kind == "default":
        return theta, None
    scaling = {k: v for k, v in rp.items() if k not in ("rope_theta", "rope_type")}
    scaling["type"] = kind
    return theta, scaling


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
        rope_theta=rope_fields(original)[0],
        rope_scaling=rope_fields(original)[1],
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


_SUPPORTED = ("llama", "granite", "gpt_bigcode", "mixtral", "granitemoe")


def _copy_tokenizer_files(src: str, dst: str) -> None:
    import shutil

    for extra in ("tokenizer.json", "tokenizer_config.json", "special_tokens_map.json", "tokenizer.model", "vocab.json",
                  "merges.txt"):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(dst, extra))


def import_from_huggingface(pretrained_model_name_or_path: str, save_path: str) -> None:
    if not os.path.isdir(pretrained_model_name_or_path):
        raise FileNotFoundError(
            f"{pretrained_model_name_or_path}: only local checkpoints can be imported (no hub access in this environment)")
    with open(os.path.join(pretrained_model_name_or_path, "config.json")) as f:
        original = json.load(f)
    model_type = original.get("model_type")
    if model_type not in _SUPPORTED:
        raise NotImplementedError(f"the current model_type ({model_type}) is not yet supported")
    m = SafeTensorsWeightsManager(pretrained_model_name_or_path)
    if model_type == "gpt_bigcode":
        config = F.import_config_bigcode(original)
        sd = F.copy_state_dict(m, config.tie_word_embeddings)
    elif model_type in ("mixtral", "granitemoe"):
        config = F.import_config_moe(original, _head_type(original))
        sd = F.import_state_dict_moe(m, config, model_type, interleave_query_key_value_tensor_for_attention)
    else:
        config = _import_config(original)
        sd = _import_state_dict(m, config)
    os.makedirs(save_path, exist_ok=True)
    SafeTensorsWeightsManager.save_state_dict(sd, save_path)
    config.save_pretrained(save_path)
    _copy_tokenizer_files(pretrained_model_name_or_path, save_path)


# ------------------------------------------------------------------------------------------------
# export: dolomite safetensors + config  ->  HuggingFace Llama / Granite layout (weight re-layout only)
# model_conversion/__init__.py:30-46, llama.py:152-290, granite.py:83-150
# ------------------------------------------------------------------------------------------------
def _export_config(config: GPTDolomiteConfig, model_type: str) -> dict:
    assert config.activation_function == "swiglu" and config.normalization_function == "rmsnorm"
    assert config.position_embedding_type == "rope"
    out = dict(
        model_type=model_type,
        architectures=["LlamaForCausalLM" if model_type == "llama" else "GraniteForCausalLM"],
        vocab_size=config.vocab_size,
        max_position_embeddings=config.n_positions,
        hidden_size=config.n_embd,
        num_hidden_layers=config.n_layer,
        num_attention_heads=config.n_head,
        num_key_value_heads=config.num_key_value_heads,
        intermediate_size=4 * config.n_embd if config.n_inner is None else config.n_inner,
        hidden_act="silu",
        rms_norm_eps=config.layer_norm_epsilon,
        use_cache=config.use_cache,
        attention_bias=config.add_bias,
        mlp_bias=config.add_bias,
        tie_word_embeddings=config.tie_word_embeddings,
        initializer_range=config.initializer_range,
        rope_theta=config.rope_theta,
        rope_scaling=config.rope_scaling,
        attention_dropout=config.attn_pdrop,
        bos_token_id=config.bos_token_id,
        eos_token_id=config.eos_token_id,
        pad_token_id=config.pad_token_id,
        torch_dtype="float32",
    )
    if model_type == "llama":
        # a Llama checkpoint has no place for the muP multipliers (llama.py:183-186)
        assert config.m_emb is None and config.m_residual is None and config.m_width is None
        assert config.attention_multiplier is None
    else:
        one = lambda x: 1 if x is None else x  # noqa: E731
        out.update(embedding_multiplier=one(config.m_emb), residual_multiplier=one(config.m_residual),
                   logits_scaling=one(config.m_width),
                   attention_multiplier=(config.n_embd // config.n_head) ** -0.5 if config.attention_multiplier is None
                   else config.attention_multiplier)
    return out


def _export_state_dict(m: SafeTensorsWeightsManager, config: GPTDolomiteConfig) -> dict:
    hd = config.n_embd // config.n_head
    sd = {
        "model.embed_tokens.weight": m.get_tensor("transformer.wte.weight"),
        "model.norm.weight": m.get_tensor("transformer.ln_f.weight"),
    }
    if m.has_tensor("lm_head.weight"):
        sd["lm_head.weight"] = m.get_tensor("lm_head.weight")
    for i in range(config.n_layer):
        src, dst = f"transformer.h.{i}.", f"model.layers.{i}."
        sd[dst + "input_layernorm.weight"] = m.get_tensor(src + "ln_1.weight")
        sd[dst + "post_attention_layernorm.weight"] = m.get_tensor(src + "ln_2.weight")
        for suffix in ("weight", "bias"):
            if not m.has_tensor(src + f"mlp.c_fc.{suffix}"):
                continue
            up, gate = split_up_gate_tensor_for_mlp(m.get_tensor(src + f"mlp.c_fc.{suffix}"))
            sd[dst + f"mlp.up_proj.{suffix}"], sd[dst + f"mlp.gate_proj.{suffix}"] = up.contiguous(), gate.contiguous()
            sd[dst + f"mlp.down_proj.{suffix}"] = m.get_tensor(src + f"mlp.c_proj.{suffix}")
            q, k, v = split_query_key_value_tensor_for_attention(
                m.get_tensor(src + f"attn.c_attn.{suffix}"), config.n_head, config.num_key_value_heads, hd,
                config.attention_head_type)
            sd[dst + f"self_attn.q_proj.{suffix}"] = q.contiguous()
            sd[dst + f"self_attn.k_proj.{suffix}"] = k.contiguous()
            sd[dst + f"self_attn.v_proj.{suffix}"] = v.contiguous()
            sd[dst + f"self_attn.o_proj.{suffix}"] = m.get_tensor(src + f"attn.c_proj.{suffix}")
    return sd


def export_to_huggingface(pretrained_model_name_or_path: str, save_path: str, model_type: str) -> None:
    """model_conversion/__init__.py:39-46: `model_type` in {"llama", "granite", "gpt_bigcode", "mixtral", "granitemoe"}"""
    if model_type == "bigcode":
        model_type = "gpt_bigcode"  # the reference's registry key (model_conversion/__init__.py:30-36)
    if model_type not in _SUPPORTED:
        raise NotImplementedError(f"the current model_type ({model_type}) is not yet supported")
    config = CommonConfig.from_pretrained(pretrained_model_name_or_path)
    m = SafeTensorsWeightsManager(pretrained_model_name_or_path)
    if model_type == "gpt_bigcode":
        out_cfg, sd = F.export_config_bigcode(config), F.copy_state_dict(m, config.tie_word_embeddings)
    elif model_type in ("mixtral", "granitemoe"):
        assert config.model_type == "moe_dolomite", f"{model_type} export needs a moe_dolomite checkpoint"
        out_cfg = F.export_config_moe(config, model_type)
        sd = F.export_state_dict_moe(m, config, model_type, split_query_key_value_tensor_for_attention)
    else:
        out_cfg, sd = _export_config(config, model_type), _export_state_dict(m, config)
    os.makedirs(save_path, exist_ok=True)
    SafeTensorsWeightsManager.save_state_dict(sd, save_path)
    with open(os.path.join(save_path, "config.json"), "w") as f:
        json.dump(out_cfg, f, indent=2)
    _copy_tokenizer_files(pretrained_model_name_or_path, save_path)
