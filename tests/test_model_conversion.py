"""CPU: import_from_huggingface (llama / granite) -- weight re-layout is exact, and the imported checkpoint evaluated by
the oracle reproduces the logits of HuggingFace's own Llama implementation on the original checkpoint
(the reference pins this in tests/hf_models/single_gpu/model_conversion_test.py)."""

import json
import os

import numpy as np
import pytest
import torch

import oracle.dolomite_oracle as O
from dolomite_engine_b200.hf_models import import_from_huggingface
from dolomite_engine_b200.hf_models.config import CommonConfig
from dolomite_engine_b200.hf_models.model_conversion import (
    interleave_query_key_value_tensor_for_attention,
    split_query_key_value_tensor_for_attention,
)
from dolomite_engine_b200.utils.safetensors import SafeTensorsWeightsManager


@pytest.mark.parametrize("head_type,nh,nkv", [("mha", 4, 4), ("gqa", 8, 2), ("mqa", 4, 1)])
def test_interleave_split_roundtrip_exact(head_type, nh, nkv):
    """tests/hf_models/single_gpu/weight_test.py:16-68"""
    hd = 8
    q, k, v = torch.randn(nh * hd, 32), torch.randn(nkv * hd, 32), torch.randn(nkv * hd, 32)
    w = interleave_query_key_value_tensor_for_attention(q, k, v, nh, nkv, hd, head_type)
    q2, k2, v2 = split_query_key_value_tensor_for_attention(w, nh, nkv, hd, head_type)
    assert torch.equal(q, q2) and torch.equal(k, k2) and torch.equal(v, v2)
    cfg = O.OracleConfig(n_embd=nh * hd, n_head=nh, num_key_value_heads=nkv, attention_head_type=head_type)
    assert torch.equal(w, O.interleave_qkv(q, k, v, cfg))


def _write_llama(path, model_type="llama", **extra):
    cfg = dict(model_type=model_type, vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
               num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128, rms_norm_eps=1e-5,
               hidden_act="silu", rope_theta=10000.0, tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
               bos_token_id=1, eos_token_id=2, pad_token_id=None, initializer_range=0.02)
    cfg.update(extra)
    g = torch.Generator().manual_seed(0)
    H, F, V, hd, nkv = 64, 128, 320, 16, 2
    sd = {"model.embed_tokens.weight": torch.randn(V, H, generator=g) * 0.05,
          "model.norm.weight": 1 + 0.1 * torch.randn(H, generator=g), "lm_head.weight": torch.randn(V, H, generator=g) * 0.05}
    for i in range(2):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"] = 1 + 0.1 * torch.randn(H, generator=g)
        sd[p + "post_attention_layernorm.weight"] = 1 + 0.1 * torch.randn(H, generator=g)
        sd[p + "self_attn.q_proj.weight"] = torch.randn(H, H, generator=g) * 0.05
        sd[p + "self_attn.k_proj.weight"] = torch.randn(nkv * hd, H, generator=g) * 0.05
        sd[p + "self_attn.v_proj.weight"] = torch.randn(nkv * hd, H, generator=g) * 0.05
        sd[p + "self_attn.o_proj.weight"] = torch.randn(H, H, generator=g) * 0.05
        sd[p + "mlp.up_proj.weight"] = torch.randn(F, H, generator=g) * 0.05
        sd[p + "mlp.gate_proj.weight"] = torch.randn(F, H, generator=g) * 0.05
        sd[p + "mlp.down_proj.weight"] = torch.randn(H, F, generator=g) * 0.05
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    SafeTensorsWeightsManager.save_state_dict(sd, path)
    return cfg, sd


def test_import_llama_matches_huggingface_llama(tmp_path):
    src, dst = str(tmp_path / "hf"), str(tmp_path / "dolomite")
    cfg, sd = _write_llama(src)
    import_from_huggingface(src, dst)
    c = CommonConfig.from_pretrained(dst)
    assert c.model_type == "gpt_dolomite" and c.attention_head_type == "gqa" and c.n_inner == 128 and not c.add_bias
    imported = SafeTensorsWeightsManager(dst).state_dict()
    up, gate = imported["transformer.h.0.mlp.c_fc.weight"].chunk(2)
    assert torch.equal(up, sd["model.layers.0.mlp.up_proj.weight"]) and torch.equal(gate, sd["model.layers.0.mlp.gate_proj.weight"])
    ocfg = O.OracleConfig(vocab_size=320, n_positions=128, n_embd=64, n_layer=2, n_head=4, num_key_value_heads=2,
                          n_inner=128, attention_head_type="gqa", tie_word_embeddings=False)
    ids = np.random.default_rng(0).integers(0, 320, size=40)
    logits = O.forward_logits(imported, ocfg, ids, np.arange(40), np.array([0, 40], dtype=np.int32))
    transformers = pytest.importorskip("transformers")
    hf_cfg = transformers.LlamaConfig(**{k: v for k, v in cfg.items() if k != "model_type"})
    hf = transformers.LlamaForCausalLM(hf_cfg).eval()
    hf.load_state_dict(sd)
    with torch.no_grad():
        ref = hf(torch.from_numpy(ids)[None]).logits[0]
    assert torch.allclose(logits, ref, atol=2e-5), (logits - ref).abs().max()


def test_import_granite_multipliers(tmp_path):
    src, dst = str(tmp_path / "hf"), str(tmp_path / "dolomite")
    _write_llama(src, model_type="granite", embedding_multiplier=12.0, residual_multiplier=0.22, logits_scaling=8.0,
                 attention_multiplier=0.0078125)
    import_from_huggingface(src, dst)
    c = CommonConfig.from_pretrained(dst)
    assert (c.m_emb, c.m_residual, c.m_width, c.attention_multiplier) == (12.0, 0.22, 8.0, 0.0078125)


def test_unsupported_family_raises(tmp_path):
    src = str(tmp_path / "x")
    os.makedirs(src)
    json.dump({"model_type": "falcon"}, open(os.path.join(src, "config.json"), "w"))
    with pytest.raises(NotImplementedError):
        import_from_huggingface(src, str(tmp_path / "y"))


@pytest.mark.parametrize("model_type", ["llama", "granite"])
def test_export_to_huggingface_is_the_exact_inverse_of_import(tmp_path, model_type):
    """model_conversion/__init__.py:39-46, llama.py:152-290: HF -> dolomite -> HF reproduces every tensor bit for bit, the
    exported config carries the HF names, and HuggingFace's own LlamaForCausalLM loads the exported directory."""
    from dolomite_engine_b200.hf_models.model_conversion import export_to_huggingface

    extra = dict(embedding_multiplier=12.0, residual_multiplier=0.22, logits_scaling=8.0, attention_multiplier=0.015625) \
        if model_type == "granite" else {}
    src, mid, dst = str(tmp_path / "hf"), str(tmp_path / "dolomite"), str(tmp_path / "hf_again")
    cfg, sd = _write_llama(src, model_type, **extra)
    import_from_huggingface(src, mid)
    export_to_huggingface(mid, dst, model_type)
    back = SafeTensorsWeightsManager(dst).state_dict()
    assert set(back) == set(sd)
    for k, v in sd.items():
        assert torch.equal(back[k], v), k
    out_cfg = json.load(open(os.path.join(dst, "config.json")))
    for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
              "num_key_value_heads", "max_position_embeddings", "rms_norm_eps", "rope_theta", "tie_word_embeddings",
              "attention_bias", "mlp_bias", "bos_token_id", "eos_token_id", *extra):
        assert out_cfg[k] == cfg[k], k
    assert out_cfg["model_type"] == model_type
    with pytest.raises(NotImplementedError):
        export_to_huggingface(mid, dst, "falcon")
    if model_type == "llama":
        from transformers import LlamaForCausalLM

        hf = LlamaForCausalLM.from_pretrained(dst, torch_dtype=torch.float32)
        assert torch.equal(hf.model.layers[1].self_attn.k_proj.weight.data, sd["model.layers.1.self_attn.k_proj.weight"])


# ---- the other families of the reference's registry: gpt_bigcode, mixtral, granitemoe --------------------------------------
def _oracle_config_from(c) -> "O.OracleConfig":
    names = set(O.OracleConfig.__dataclass_fields__) - {"extra"}
    kw = {k: getattr(c, k) for k in names if hasattr(c, k) and k not in ("num_experts", "num_experts_per_tok")}
    if c.model_type == "moe_dolomite":
        kw.update(num_experts=c.num_experts, num_experts_per_tok=c.num_experts_per_tok)
    return O.OracleConfig(**kw)


def _tiny_hf(family):
    import transformers as T

    torch.manual_seed(0)
    if family == "gpt_bigcode":
        cfg = T.GPTBigCodeConfig(vocab_size=320, n_embd=64, n_layer=2, n_head=4, n_positions=128, n_inner=128, multi_query=True,
                                 activation_function="gelu_pytorch_tanh", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                                 bos_token_id=1, eos_token_id=2, pad_token_id=None)
        return T.GPTBigCodeForCausalLM(cfg).eval()
    common = dict(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                  num_key_value_heads=2, max_position_embeddings=128, num_local_experts=8, num_experts_per_tok=2,
                  tie_word_embeddings=False, bos_token_id=1, eos_token_id=2, pad_token_id=None, rms_norm_eps=1e-5)
    if family == "mixtral":
        return T.MixtralForCausalLM(T.MixtralConfig(sliding_window=None, **common)).eval()
    return T.GraniteMoeForCausalLM(T.GraniteMoeConfig(embedding_multiplier=6.0, residual_multiplier=0.5, logits_scaling=4.0,
                                                       attention_multiplier=0.125, **common)).eval()


@pytest.mark.parametrize("family", ["gpt_bigcode", "mixtral", "granitemoe"])
def test_more_families_import_matches_huggingface_and_export_is_the_inverse(tmp_path, family):
    """model_conversion/{bigcode,mixtral,granitemoe}.py: (1) the imported checkpoint, evaluated by the oracle, reproduces the
    logits of HuggingFace's own implementation; (2) HF -> dolomite -> HF is bit-identical and HuggingFace loads it"""
    import transformers as T

    from dolomite_engine_b200.hf_models.model_conversion import export_to_huggingface

    hf = _tiny_hf(family)
    with torch.no_grad():  # make every parameter non-trivial (norm weights / biases are initialised to 1 / 0)
        for p in hf.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    src, mid, dst = str(tmp_path / "hf"), str(tmp_path / "dolomite"), str(tmp_path / "hf_again")
    hf.save_pretrained(src, safe_serialization=True)
    import_from_huggingface(src, mid)
    c = CommonConfig.from_pretrained(mid)
    assert c.model_type == ("gpt_dolomite" if family == "gpt_bigcode" else "moe_dolomite")
    imported = SafeTensorsWeightsManager(mid).state_dict()
    ids = np.random.default_rng(1).integers(3, 320, size=24)
    logits = O.forward_logits(imported, _oracle_config_from(c), ids, np.arange(24), np.array([0, 24], dtype=np.int32))
    with torch.no_grad():
        ref = hf(torch.from_numpy(ids)[None]).logits[0]
    assert torch.allclose(logits, ref, atol=5e-5), (logits - ref).abs().max()
    if family == "granitemoe":
        assert (c.m_emb, c.m_residual, c.m_width, c.attention_multiplier) == (6.0, 0.5, 4.0, 0.125)
    if family != "gpt_bigcode":
        assert c.num_experts == 8 and c.num_experts_per_tok == 2
        assert imported["transformer.h.0.mlp.c_fc.weight"].shape == (8, 256, 64)
    # export: the exact inverse, loadable by HuggingFace
    export_to_huggingface(mid, dst, family)
    original = SafeTensorsWeightsManager(src).state_dict()
    back = SafeTensorsWeightsManager(dst).state_dict()
    tied = {"lm_head.weight"} if family == "gpt_bigcode" else set()
    assert set(back) | tied == set(original) | tied
    for k, v in back.items():
        assert torch.equal(v, original[k]), k
    again = T.AutoModelForCausalLM.from_pretrained(dst, torch_dtype=torch.float32).eval()
    with torch.no_grad():
        assert torch.equal(again(torch.from_numpy(ids)[None]).logits[0], ref)


def test_rope_fields_reads_both_config_dialects(tmp_path):
    """transformers >= 5 writes `rope_parameters` instead of `rope_theta` / `rope_scaling`; Llama-3's theta (500000) must
    survive the import either way"""
    from dolomite_engine_b200.hf_models.model_conversion import rope_fields

    assert rope_fields({"rope_theta": 500000.0}) == (500000.0, None)
    assert rope_fields({}) == (10000, None)
    assert rope_fields({"rope_parameters": {"rope_theta": 500000.0, "rope_type": "default"}}) == (500000.0, None)
    theta, scaling = rope_fields({"rope_parameters": {"rope_theta": 1e6, "rope_type": "yarn", "factor": 4.0,
                                                      "original_max_position_embeddings": 128}})
    assert theta == 1e6 and scaling == {"type": "yarn", "factor": 4.0, "original_max_position_embeddings": 128}
    src, dst = str(tmp_path / "hf"), str(tmp_path / "dolomite")
    cfg, _ = _write_llama(src)
    cfg.pop("rope_theta")
    cfg["rope_parameters"] = {"rope_theta": 500000.0, "rope_type": "default"}
    json.dump(cfg, open(os.path.join(src, "config.json"), "w"))
    import_from_huggingface(src, dst)
    c = CommonConfig.from_pretrained(dst)
    assert c.rope_theta == 500000.0 and c.rope_scaling is None
