"""Training-mode dropout on the B200 path against the CPU oracle with the SAME masks.

The reference draws its masks from torch's Philox stream (nn.Dropout at gpt_dolomite/base.py:138, attention/base.py:91-92,
gpt_dolomite/mlp.py:43; `dropout_p` of flash_attn_varlen_func at attention/padding_free.py:49-59), which no independent
kernel reproduces, so parity is stated as: (a) the masks are a documented integer function (csrc/common.cuh) that the oracle
restates in numpy bit for bit -- checked element by element here; (b) with those masks installed, the oracle's loss and every
gradient match the GPU's to the usual tolerances; (c) keep rate 1 - p and scale 1 / (1 - p); (d) backward and recomputed
(checkpointed) blocks see the masks of their forward; (e) eval mode is the identity."""

import math

import numpy as np
import pytest
import torch

import oracle.dolomite_oracle as O

pytestmark = pytest.mark.gpu


def K():
    from dolomite_engine_b200 import kernels

    return kernels


def bf(x):
    return x.to(torch.bfloat16)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("p,post_mul,with_res", [(0.1, 1.0, True), (0.25, 0.22, True), (0.5, 1.0, False), (0.0, 0.3, False)])
def test_elementwise_dropout_is_bit_exact_with_the_oracle_masks(p, post_mul, with_res):
    g = torch.Generator().manual_seed(3)
    T, H = 333, 520
    x, res, dy = bf(torch.randn(T, H, generator=g)), bf(torch.randn(T, H, generator=g)), bf(torch.randn(T, H, generator=g))
    d = O.DropoutOracle(4242)
    keys = K().dropout_keys(4242, 9)
    assert keys == d.keys(9)
    scale = d.flat_scale(9, (T, H), p)
    y = K().dropout_fwd(x.cuda(), p, keys, residual=res.cuda() if with_res else None, post_mul=post_mul)
    ref = bf(x.float() * scale)
    if post_mul != 1.0:
        ref = bf(ref.float() * post_mul)
    if with_res:
        ref = bf(res.float() + ref.float())
    assert torch.equal(y.cpu(), ref)
    dx = K().dropout_bwd(dy.cuda(), p, keys, pre_mul=post_mul)
    gref = dy if post_mul == 1.0 else bf(dy.float() * post_mul)
    assert torch.equal(dx.cpu(), bf(gref.float() * scale))
    if p > 0:
        kept = (scale > 0).float().mean().item()
        assert abs(kept - (1 - p)) < 4 * math.sqrt(p * (1 - p) / (T * H)) + 1e-3


def _attn_inputs(lens, ng, g, hd, seed=7):
    gen = torch.Generator().manual_seed(seed)
    T = sum(lens)
    qkv = bf(torch.randn(T, ng * (g + 2) * hd, generator=gen))
    dout = bf(torch.randn(T, ng * g * hd, generator=gen))
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return qkv, dout, cu


@pytest.mark.parametrize("lens,ng,g,hd", [([200, 130, 315], 2, 1, 80), ([100, 37, 300, 1, 129], 2, 1, 64), ([300, 77, 260], 1, 2, 128),
                                          ([150, 250], 1, 2, 32), ([260], 2, 1, 96)])
def test_attention_dropout_fwd_bwd_vs_oracle(lens, ng, g, hd):
    """every attention kernel (single-buffer / split forward, pipelined / serial backward) with dropout_p = 0.2"""
    p_drop, seed, site = 0.2, 777, 7
    qkv, dout, cu = _attn_inputs(lens, ng, g, hd)
    scale = 1.0 / math.sqrt(hd)
    cfg = O.OracleConfig(n_embd=ng * g * hd, n_head=ng * g, num_key_value_heads=ng,
                         attention_head_type="mha" if g == 1 else "gqa")
    x = qkv.float().requires_grad_(True)
    q, k, v = O.split_qkv_activations(x, cfg)
    O.DROPOUT = O.DropoutOracle(seed)
    try:
        ref = O.packed_causal_attention(q, k, v, cu, scale, dropout_site=site, dropout_p=p_drop)
        ref.backward(dout.float())
        ref_nodrop = O.packed_causal_attention(q.detach(), k.detach(), v.detach(), cu, scale)
    finally:
        O.DROPOUT = None
    keys = K().dropout_keys(seed, site)
    args = (torch.from_numpy(cu).cuda(), max(lens), ng, g, hd, scale)
    outs = []
    default_split = K().get_option("attn_fwd_split")
    splits = (default_split, 0, 2, 3) if hd >= 64 else (default_split,)
    for split in splits:
        try:
            K().set_option("attn_fwd_split", split)
            out, lse = K().attn_varlen_fwd(qkv.cuda(), *args, dropout_p=p_drop, dropout_keys=keys)
        finally:
            K().set_option("attn_fwd_split", default_split)
        assert rel_l2(out, ref) < 8e-3, split
        assert rel_l2(out, ref_nodrop) > 0.2  # the masks did something
        outs.append((out, lse))
    out, lse = outs[0]
    # the log-sum-exp is that of the UNdropped row
    _, lse0 = K().attn_varlen_fwd(qkv.cuda(), *args)
    assert torch.allclose(lse, lse0, atol=2e-3, rtol=1e-4)
    dqkv = K().attn_varlen_bwd(dout.cuda(), qkv.cuda(), out, lse, *args, dropout_p=p_drop, dropout_keys=keys)
    assert rel_l2(dqkv, x.grad) < 1.5e-2
    # other keys -> other masks
    out2, _ = K().attn_varlen_fwd(qkv.cuda(), *args, dropout_p=p_drop, dropout_keys=K().dropout_keys(seed + 1, site))
    assert rel_l2(out2, out) > 0.2


DROP_CONFIGS = {
    "hd80_bias_mup": dict(vocab_size=1024, n_positions=512, n_embd=320, n_layer=2, n_head=4, n_inner=640, attention_head_type="mha",
                          add_bias=True, m_emb=12.0, m_residual=0.22, m_width=2.0),
    "hd128_gqa": dict(vocab_size=1024, n_positions=512, n_embd=512, n_layer=1, n_head=4, num_key_value_heads=2, n_inner=1024,
                      attention_head_type="gqa", add_bias=False, tie_word_embeddings=False),
    "bigcode": dict(vocab_size=1024, n_positions=512, n_embd=256, n_layer=2, n_head=4, n_inner=1024, attention_head_type="mqa",
                    add_bias=True, position_embedding_type="learned_absolute", normalization_function="layernorm",
                    activation_function="gelu_pytorch_tanh", m_emb=3.0),
}
PDROP = dict(resid_pdrop=0.1, embd_pdrop=0.15, attn_pdrop=0.2)


def _build(name, **extra):
    from dolomite_engine_b200.hf_models import GPTDolomiteConfig, GPTDolomiteForCausalLM

    kw = dict(DROP_CONFIGS[name])
    ocfg = O.OracleConfig(**kw, **PDROP)
    params = O.init_params(ocfg, seed=42)
    g = torch.Generator().manual_seed(7)
    for k_ in params:
        if k_.endswith(".bias"):
            params[k_] = torch.randn(params[k_].shape, generator=g) * 0.02
    d = dict(position_embedding_type="rope", normalization_function="rmsnorm", activation_function="swiglu", eos_token_id=7)
    d.update(kw)
    d.update(PDROP)
    d.update(extra)
    model = GPTDolomiteForCausalLM(GPTDolomiteConfig(**d), seed=None)
    model.load_state_dict(params)
    return model, ocfg, params


def _batch(ocfg, ragged=True, mbs=2, seq=160, seed=1234):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, ocfg.vocab_size, size=(mbs, seq + 1), dtype=np.int64)
    t[0, 20] = 7
    t[1, 5] = 7
    t[1, 140] = 7
    inp, labels = O.split_tokens(t)
    b = O.prepare_model_inputs(inp.copy(), 7, ragged, ragged)
    args = (torch.from_numpy(b["input_ids"]).cuda(), torch.from_numpy(b["position_ids"]).cuda(),
            torch.from_numpy(b["cu_seqlens"]).cuda(), b["max_seqlen"],
            torch.from_numpy(np.ascontiguousarray(labels).reshape(-1)).cuda())
    return t, args


@pytest.mark.parametrize("name", list(DROP_CONFIGS))
def test_model_with_dropout_matches_the_oracle_with_the_same_masks(name):
    model, ocfg, params = _build(name)
    model.assume_unit_loss_grad = True
    tokens, args = _batch(ocfg)
    eng = model.engine
    eng.dropout_seed = 31337
    model.train()
    eng.zero_grad()
    loss = model.forward_pretraining_loss(*args)
    assert eng._saved["dropout_seed"] == 31337
    loss.backward()
    torch.cuda.synchronize()
    p_req = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.DROPOUT = O.DropoutOracle(31337)
    try:
        ref, _ = O.pretraining_loss(p_req, ocfg, tokens, 7, True, True)
        ref.backward()
    finally:
        O.DROPOUT = None
    ref_eval, _ = O.pretraining_loss(params, ocfg, tokens, 7, True, True)
    assert abs(loss.item() - ref.item()) / ref.item() < 1.5e-3, (loss.item(), ref.item(), ref_eval.item())
    p_ev = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.pretraining_loss(p_ev, ocfg, tokens, 7, True, True)[0].backward()
    moved = 0
    for n, u, _ in eng.named_views():
        r = p_req[n].grad
        if r is None or r.norm() == 0:
            continue
        assert rel_l2(u.gviews[n], r) < 4e-2, (n, rel_l2(u.gviews[n], r))
        moved += int(rel_l2(p_ev[n].grad, r) > 0.1)
    assert moved >= 4  # the masks moved the gradients far beyond the tolerance: the comparison above is not vacuous
    # the next pass draws new masks (seed + 1) ...
    eng.zero_grad()
    loss2 = model.forward_pretraining_loss(*args)
    assert eng._saved["dropout_seed"] == 31338 and loss2.item() != loss.item()
    # ... and evaluation mode is the identity
    model.eval()
    with torch.no_grad():
        loss_eval = model.forward_pretraining_loss(*args)
    assert abs(loss_eval.item() - ref_eval.item()) / ref_eval.item() < 1e-3
    model.train()


def test_checkpointed_blocks_regenerate_the_masks_of_their_forward():
    model, ocfg, _ = _build("hd80_bias_mup")
    model.assume_unit_loss_grad = True
    _, args = _batch(ocfg)
    eng = model.engine
    grads = []
    for ck in (None, 1):
        eng.checkpoint_every = ck
        eng.dropout_seed, eng._dropout_passes = 555, 0
        eng.zero_grad()
        loss = model.forward_pretraining_loss(*args)
        loss.backward()
        torch.cuda.synchronize()
        grads.append((loss.item(), {n: u.gviews[n].clone() for n, u, _ in eng.named_views()}))
    eng.checkpoint_every = None
    assert grads[0][0] == grads[1][0]
    for n, g in grads[0][1].items():
        assert rel_l2(grads[1][1][n], g) < 1e-5, n


def test_moe_block_with_dropout_matches_the_oracle():
    from dolomite_engine_b200.hf_models import MoEDolomiteConfig, MoEDolomiteForCausalLM

    kw = dict(vocab_size=1024, n_positions=512, n_embd=256, n_layer=2, n_head=4, n_inner=256, attention_head_type="mha", add_bias=False,
              num_experts=8, num_experts_per_tok=2, m_residual=0.5)
    ocfg = O.OracleConfig(**kw, **PDROP)
    params = O.init_params(ocfg, seed=42)
    model = MoEDolomiteForCausalLM(MoEDolomiteConfig(position_embedding_type="rope", normalization_function="rmsnorm",
                                                     activation_function="swiglu", eos_token_id=7, **kw, **PDROP), seed=None)
    model.load_state_dict(params)
    model.assume_unit_loss_grad = True
    tokens, args = _batch(ocfg)
    eng = model.engine
    eng.dropout_seed = 99
    eng.zero_grad()
    loss = model.forward_pretraining_loss(*args)
    loss.backward()
    torch.cuda.synchronize()
    O.DROPOUT = O.DropoutOracle(99)
    try:
        ref, _ = O.pretraining_loss(params, ocfg, tokens, 7, True, True)
    finally:
        O.DROPOUT = None
    # free routing: a handful of bf16 near-ties may pick another expert (tests/test_gpu_moe.py), hence the looser bar
    assert abs(loss.item() - ref.item()) / ref.item() < 5e-3, (loss.item(), ref.item())
