"""CPU: the oracle reproduces the golden vectors generated from the reference's leaf modules
(oracle/validate_against_reference.py).  These fixtures are the pin that lets the GPU box check parity without
/root/reference."""

import os

import numpy as np
import pytest
import torch

import oracle.dolomite_oracle as O
from oracle.validate_against_reference import CONFIGS


def _params(cfg):
    p = O.init_params(cfg, seed=42)
    if cfg.add_bias:
        g = torch.Generator().manual_seed(7)
        for k in p:
            if k.endswith(".bias"):
                p[k] = torch.randn(p[k].shape, generator=g) * 0.02
    return p


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("mode", ["uniform", "ragged"])
def test_model_against_golden(golden_dir, name, mode):
    cfg = O.OracleConfig(**CONFIGS[name])
    fx = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    p = {k: v.clone().requires_grad_(True) for k, v in _params(cfg).items()}
    ram = rpi = mode == "ragged"
    tokens = fx["tokens"]
    inp, _ = O.split_tokens(tokens)
    b = O.prepare_model_inputs(inp.copy(), int(fx["eos"]), ram, rpi)
    # integer bookkeeping: bit exact
    assert np.array_equal(b["cu_seqlens"], fx[f"{mode}_cu_seqlens"]) and b["cu_seqlens"].dtype == np.int32
    assert np.array_equal(b["position_ids"], fx[f"{mode}_position_ids"])
    assert b["max_seqlen"] == int(fx[f"{mode}_max_seqlen"])
    loss, logits = O.pretraining_loss(p, cfg, tokens, int(fx["eos"]), ram, rpi)
    loss.backward()
    np.testing.assert_allclose(loss.item(), fx[f"{mode}_loss"], rtol=1e-6)
    np.testing.assert_allclose(logits.detach()[::8].numpy(), fx[f"{mode}_logits_rows"], atol=3e-5)
    np.testing.assert_allclose(p["transformer.ln_f.weight"].grad.numpy(), fx[f"{mode}_grad_ln_f"], atol=1e-6, rtol=1e-4)
    np.testing.assert_allclose(p["transformer.h.0.attn.c_attn.weight"].grad[::4].numpy(), fx[f"{mode}_grad_c_attn_0"],
                               atol=1e-6, rtol=1e-4)


def test_moe_against_golden(golden_dir):
    fx = np.load(os.path.join(golden_dir, "moe_layer.npz"))
    cfg = O.OracleConfig(vocab_size=256, n_embd=64, n_layer=1, n_head=4, n_inner=128, num_experts=8, num_experts_per_tok=2)
    p = {"m.gate.weight": torch.from_numpy(fx["gate"]), "m.c_fc.weight": torch.from_numpy(fx["c_fc"]),
         "m.c_proj.weight": torch.from_numpy(fx["c_proj"])}
    x = torch.from_numpy(fx["x"])
    y, logits = O.sparse_moe(x, p, "m.", cfg)
    np.testing.assert_allclose(y.numpy(), fx["y"], atol=1e-6)
    np.testing.assert_allclose(logits.numpy(), fx["router_logits"], atol=1e-6)
    _, idx, _ = O.moe_route(x, p["m.gate.weight"], 2)
    assert np.array_equal(O.moe_expert_counts(idx, 8), fx["counts"])  # bit exact


def test_bookkeeping_against_golden(golden_dir):
    fx = np.load(os.path.join(golden_dir, "bookkeeping.npz"))
    ids = [[5, 6, 7, 8, 9, 1, 2, 3, 4, 5], [9, 8, 7, 6, 5]]
    o = O.convert_padding_free_lists_to_tensors(ids, labels=ids)
    assert np.array_equal(o["cu_seqlens"], fx["cu_seqlens"]) and o["cu_seqlens"].dtype == np.int32
    assert o["max_seqlen"] == int(fx["max_seqlen"])
    assert np.array_equal(o["position_ids"], fx["position_ids"])
    assert np.array_equal(o["input_ids"], fx["input_ids"])
    assert np.array_equal(O.finetune_shift_labels(o["labels"], o["cu_seqlens"]), fx["shift_labels"])


def test_bf16_emulation_is_close_to_fp32():
    cfg = O.OracleConfig(**CONFIGS["c1"])
    p = _params(cfg)
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, cfg.vocab_size, size=(2, 65), dtype=np.int64)
    l32, _ = O.pretraining_loss(p, cfg, tokens)
    l16, _ = O.pretraining_loss(p, cfg, tokens, bf16=True)
    assert abs(l32.item() - l16.item()) / l32.item() < 2e-3


def test_flop_model_matches_survey_numbers():
    # C2 = 24.914 GFLOP/token, Llama-3-8B = 57.913 GFLOP/token
    c2 = O.OracleConfig(vocab_size=49152, n_embd=2560, n_layer=32, n_head=32, n_inner=10240, n_positions=4096)
    assert abs(O.model_flops_per_token(c2, 4096) / 1e9 - 24.914) < 1e-3
    l8 = O.OracleConfig(vocab_size=128256, n_embd=4096, n_layer=32, n_head=32, num_key_value_heads=8, n_inner=14336,
                        attention_head_type="gqa", n_positions=8192, tie_word_embeddings=False)
    assert abs(O.model_flops_per_token(l8, 8192) / 1e9 - 57.913) < 1e-3
