"""The other checkpoint families of the reference's `import_from_huggingface` / `export_to_huggingface`
(model_conversion/__init__.py:11-46): gpt_bigcode (bigcode.py), mixtral (mixtral.py) and granitemoe (granitemoe.py).

  gpt_bigcode   tensor names and layouts ARE the dolomite ones (transformer.h.N.attn.c_attn ...): only config.json changes
  mixtral       per-expert `block_sparse_moe.experts.E.{w1 (gate), w3 (up), w2 (down)}` <-> stacked
                `mlp.c_fc [E, 2F, H] = cat(up, gate)`, `mlp.c_proj [E, H, F]`; router `block_sparse_moe.gate` <-> `mlp.gate`
  granitemoe    fused `block_sparse_moe.input_linear [E, 2F, H] = cat(gate, up)` <-> `mlp.c_fc` with the halves swapped,
                `output_linear` <-> `mlp.c_proj`, `router.layer` <-> `mlp.gate`; muP multipliers as in granite

Attention projections use the same q/k/v interleave as llama.  Local directories only (no hub access).
"""

from __future__ import annotations

import torch

from ..utils.safetensors import SafeTensorsWeightsManager
from .config import GPTDolomiteConfig, MoEDolomiteConfig

_ONE_TO_NONE = lambda x: None if x == 1 else x  # noqa: E731
_NONE_TO_ONE = lambda x: 1 if x is None else x  # noqa: E731


# ------------------------------------------------------------------------------------------------
# gpt_bigcode (StarCoder): config only
# ------------------------------------------------------------------------------------------------
def import_config_bigcode(o: dict) -> GPTDolomiteConfig:
    act = o.get("activation_function", "gelu_pytorch_tanh")
    assert act in ("gelu_pytorch_tanh", "gelu"), act
    return GPTDolomiteConfig(
        vocab_size=o["vocab_size"], n_positions=o["n_positions"], n_embd=o["n_embd"], n_layer=o["n_layer"], n_head=o["n_head"],
        attention_head_type="mqa" if o.get("multi_query", True) else "mha", position_embedding_type="learned_absolute",
        n_inner=o.get("n_inner"), activation_function=act, normalization_function="layernorm",
        layer_norm_epsilon=o.get("layer_norm_epsilon", 1e-5), use_cache=o.get("use_cache", True), add_bias=True,
        tie_word_embeddings=o.get("tie_word_embeddings", True), initializer_range=o.get("initializer_range", 0.02),
        attn_pdrop=o.get("attn_pdrop", 0.1), resid_pdrop=o.get("resid_pdrop", 0.1), embd_pdrop=o.get("embd_pdrop", 0.1),
        scale_attn_weights=o.get("scale_attn_weights", True), attention_softmax_in_fp32=o.get("attention_softmax_in_fp32", True),
        bos_token_id=o.get("bos_token_id"), eos_token_id=o.get("eos_token_id"), pad_token_id=o.get("pad_token_id"))


def export_config_bigcode(c: GPTDolomiteConfig) -> dict:
    assert c.activation_function == "gelu_pytorch_tanh" and c.normalization_function == "layernorm"
    assert c.attention_head_type in ("mha", "mqa") and c.position_embedding_type == "learned_absolute"
    assert c.m_emb is None and c.m_residual is None and c.m_width is None and c.attention_multiplier is None
    return dict(model_type="gpt_bigcode", architectures=["GPTBigCodeForCausalLM"], vocab_size=c.vocab_size,
                n_positions=c.n_positions, n_embd=c.n_embd, n_layer=c.n_layer, n_head=c.n_head, n_inner=c.n_inner,
                activation_function=c.activation_function, resid_pdrop=c.resid_pdrop, embd_pdrop=c.embd_pdrop,
                attn_pdrop=c.attn_pdrop, layer_norm_epsilon=c.layer_norm_epsilon, initializer_range=c.initializer_range,
                scale_attn_weights=c.scale_attn_weights, use_cache=c.use_cache,
                attention_softmax_in_fp32=c.attention_softmax_in_fp32, multi_query=c.attention_head_type == "mqa",
                tie_word_embeddings=c.tie_word_embeddings, bos_token_id=c.bos_token_id, eos_token_id=c.eos_token_id,
                pad_token_id=c.pad_token_id, torch_dtype="float32")


def copy_state_dict(m: SafeTensorsWeightsManager, tied: bool) -> dict:
    sd = m.state_dict()
    if tied:
        sd.pop("lm_head.weight", None)  # a tied head is not serialised on either side
    return sd


# ------------------------------------------------------------------------------------------------
# mixtral / granitemoe: config
# ------------------------------------------------------------------------------------------------
def _rope(o: dict):
    from .model_conversion import rope_fields  # (model_conversion imports this module)

    return rope_fields(o)


def _moe_common(o: dict, head_type: str) -> dict:
    return dict(
        vocab_size=o["vocab_size"], n_positions=o["max_position_embeddings"], n_embd=o["hidden_size"],
        n_layer=o["num_hidden_layers"], n_head=o["num_attention_heads"],
        num_key_value_heads=o.get("num_key_value_heads") or o["num_attention_heads"], attention_head_type=head_type,
        position_embedding_type="rope", n_inner=o["intermediate_size"], activation_function="swiglu",
        normalization_function="rmsnorm", layer_norm_epsilon=o.get("rms_norm_eps", 1e-5), use_cache=o.get("use_cache", True),
        tie_word_embeddings=o.get("tie_word_embeddings", False), initializer_range=o.get("initializer_range", 0.02),
        rope_theta=_rope(o)[0], attn_pdrop=o.get("attention_dropout", 0.0), resid_pdrop=0.0, embd_pdrop=0.0,
        num_experts=o["num_local_experts"], num_experts_per_tok=o["num_experts_per_tok"],
        output_router_logits=o.get("output_router_logits", False), router_aux_loss_coef=o.get("router_aux_loss_coef", 0.001),
        bos_token_id=o.get("bos_token_id"), eos_token_id=o.get("eos_token_id"), pad_token_id=o.get("pad_token_id"))


def import_config_moe(o: dict, head_type: str) -> MoEDolomiteConfig:
    kw = _moe_common(o, head_type)
    if o["model_type"] == "mixtral":
        assert o.get("hidden_act", "silu") == "silu"
        kw.update(add_bias=False)
    else:  # granitemoe
        assert o.get("hidden_act", o.get("activation_function", "silu")) == "silu"
        assert not o.get("attention_bias", False), "granitemoe checkpoints with attention bias are not supported"
        kw.update(add_bias=False, rope_scaling=_rope(o)[1], m_emb=_ONE_TO_NONE(o.get("embedding_multiplier", 1)),
                  m_residual=_ONE_TO_NONE(o.get("residual_multiplier", 1)), m_width=_ONE_TO_NONE(o.get("logits_scaling", 1)),
                  attention_multiplier=o.get("attention_multiplier"))
    return MoEDolomiteConfig(**kw)


def export_config_moe(c: MoEDolomiteConfig, model_type: str) -> dict:
    assert c.activation_function == "swiglu" and c.normalization_function == "rmsnorm" and c.position_embedding_type == "rope"
    assert not c.add_bias
    out = dict(
        model_type=model_type, vocab_size=c.vocab_size, max_position_embeddings=c.n_positions, hidden_size=c.n_embd,
        num_hidden_layers=c.n_layer, num_attention_heads=c.n_head, num_key_value_heads=c.num_key_value_heads,
        intermediate_size=4 * c.n_embd if c.n_inner is None else c.n_inner, hidden_act="silu", rms_norm_eps=c.layer_norm_epsilon,
        use_cache=c.use_cache, tie_word_embeddings=c.tie_word_embeddings, initializer_range=c.initializer_range,
        rope_theta=c.rope_theta, rope_scaling=c.rope_scaling, attention_dropout=c.attn_pdrop, num_local_experts=c.num_experts,
        num_experts_per_tok=c.num_experts_per_tok, output_router_logits=c.output_router_logits,
        router_aux_loss_coef=c.router_aux_loss_coef, bos_token_id=c.bos_token_id, eos_token_id=c.eos_token_id,
        pad_token_id=c.pad_token_id, torch_dtype="float32")
    if model_type == "mixtral":
        assert c.m_emb is None and c.m_residual is None and c.m_width is None and c.attention_multiplier is None
        out["architectures"] = ["MixtralForCausalLM"]
    else:
        out.update(architectures=["GraniteMoeForCausalLM"], attention_bias=False, embedding_multiplier=_NONE_TO_ONE(c.m_emb),
                   residual_multiplier=_NONE_TO_ONE(c.m_residual), logits_scaling=_NONE_TO_ONE(c.m_width),
                   attention_multiplier=(c.n_embd // c.n_head) ** -0.5 if c.attention_multiplier is None
                   else c.attention_multiplier)
    return out


# ------------------------------------------------------------------------------------------------
# mixtral / granitemoe: tensors
# ------------------------------------------------------------------------------------------------
def _swap_halves(w: torch.Tensor) -> torch.Tensor:
    """[E, 2F, H]: cat(a, b) -> cat(b, a) along the 2F axis (granitemoe stores gate first, dolomite up first)"""
    a, b = w.chunk(2, dim=1)
    return torch.cat([b, a], dim=1).contiguous()


def import_state_dict_moe(m: SafeTensorsWeightsManager, c: MoEDolomiteConfig, family: str, interleave_qkv) -> dict:
    hd = c.n_embd // c.n_head
    sd = {"transformer.wte.weight": m.get_tensor("model.embed_tokens.weight"),
          "transformer.ln_f.weight": m.get_tensor("model.norm.weight")}
    if m.has_tensor("lm_head.weight") and not c.tie_word_embeddings:
        sd["lm_head.weight"] = m.get_tensor("lm_head.weight")
    for i in range(c.n_layer):
        src, dst = f"model.layers.{i}.", f"transformer.h.{i}."
        sd[dst + "ln_1.weight"] = m.get_tensor(src + "input_layernorm.weight")
        sd[dst + "ln_2.weight"] = m.get_tensor(src + "post_attention_layernorm.weight")
        moe = src + "block_sparse_moe."
        if family == "mixtral":
            sd[dst + "mlp.gate.weight"] = m.get_tensor(moe + "gate.weight")
            sd[dst + "mlp.c_fc.weight"] = torch.stack([
                torch.cat([m.get_tensor(f"{moe}experts.{e}.w3.weight"), m.get_tensor(f"{moe}experts.{e}.w1.weight")])
                for e in range(c.num_experts)])
            sd[dst + "mlp.c_proj.weight"] = torch.stack([m.get_tensor(f"{moe}experts.{e}.w2.weight") for e in range(c.num_experts)])
        else:
            sd[dst + "mlp.gate.weight"] = m.get_tensor(moe + "router.layer.weight")
            sd[dst + "mlp.c_fc.weight"] = _swap_halves(m.get_tensor(moe + "input_linear.weight"))
            sd[dst + "mlp.c_proj.weight"] = m.get_tensor(moe + "output_linear.weight")
        sd[dst + "attn.c_attn.weight"] = interleave_qkv(
            m.get_tensor(src + "self_attn.q_proj.weight"), m.get_tensor(src + "self_attn.k_proj.weight"),
            m.get_tensor(src + "self_attn.v_proj.weight"), c.n_head, c.num_key_value_heads, hd, c.attention_head_type)
        sd[dst + "attn.c_proj.weight"] = m.get_tensor(src + "self_attn.o_proj.weight")
    return sd


def export_state_dict_moe(m: SafeTensorsWeightsManager, c: MoEDolomiteConfig, family: str, split_qkv) -> dict:
    hd = c.n_embd // c.n_head
    sd = {"model.embed_tokens.weight": m.get_tensor("transformer.wte.weight"),
          "model.norm.weight": m.get_tensor("transformer.ln_f.weight")}
    if m.has_tensor("lm_head.weight"):
        sd["lm_head.weight"] = m.get_tensor("lm_head.weight")
    for i in range(c.n_layer):
        src, dst = f"transformer.h.{i}.", f"model.layers.{i}."
        sd[dst + "input_layernorm.weight"] = m.get_tensor(src + "ln_1.weight")
        sd[dst + "post_attention_layernorm.weight"] = m.get_tensor(src + "ln_2.weight")
        moe = dst + "block_sparse_moe."
        c_fc, c_proj = m.get_tensor(src + "mlp.c_fc.weight"), m.get_tensor(src + "mlp.c_proj.weight")
        if family == "mixtral":
            sd[moe + "gate.weight"] = m.get_tensor(src + "mlp.gate.weight")
            for e in range(c.num_experts):
                up, gate = c_fc[e].chunk(2)
                sd[f"{moe}experts.{e}.w3.weight"], sd[f"{moe}experts.{e}.w1.weight"] = up.contiguous(), gate.contiguous()
                sd[f"{moe}experts.{e}.w2.weight"] = c_proj[e].contiguous()
        else:
            sd[moe + "router.layer.weight"] = m.get_tensor(src + "mlp.gate.weight")
            sd[moe + "input_linear.weight"] = _swap_halves(c_fc)
            sd[moe + "output_linear.weight"] = c_proj
        q, k, v = split_qkv(m.get_tensor(src + "attn.c_attn.weight"), c.n_head, c.num_key_value_heads, hd, c.attention_head_type)
        sd[dst + "self_attn.q_proj.weight"], sd[dst + "self_attn.k_proj.weight"] = q.contiguous(), k.contiguous()
        sd[dst + "self_attn.v_proj.weight"] = v.contiguous()
        sd[dst + "self_attn.o_proj.weight"] = m.get_tensor(src + "attn.c_proj.weight")
    return sd
