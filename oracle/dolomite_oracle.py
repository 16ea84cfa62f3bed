"""CPU oracle: a plain restatement of the reference algorithm for the GPTDolomite / MoEDolomite training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under dolomite_engine_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker or the
timed CPU baseline, never as the product path.

Parity status: PINNED against the reference's own leaf modules (GPTDolomiteBlock, Attention, MLP, RMSNorm, RoPE,
SparseMoE, the QKV/up-gate (de)interleave helpers) imported read-only from /root/reference by
oracle/validate_against_reference.py, which also writes the golden fixtures under tests/golden/.  The reference
stores no golden vectors of its own; its tests are cross-implementation equivalences.

Every function cites the reference file:line (relative to dolomite_engine/) it restates.  Float math is torch on
the CPU in fp32 (optionally emulating the bf16 rounding points of the mixed-precision path); integer bookkeeping
is numpy and must match bit-exactly.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# config (hf_models/config.py:6-111) -- only the fields the hot path reads
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    vocab_size: int = 2048
    n_positions: int = 1024
    n_embd: int = 256
    n_layer: int = 2
    n_head: int = 4
    num_key_value_heads: int | None = None
    n_inner: int | None = None
    activation_function: str = "swiglu"
    attention_head_type: str = "mha"
    normalization_function: str = "rmsnorm"
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    scale_attn_weights: bool = True
    attention_multiplier: float | None = None
    add_bias: bool = False
    position_embedding_type: str = "rope"
    rope_theta: float = 10000
    m_emb: float | None = None
    m_width: float | None = None
    m_residual: float | None = None
    init_method: str = "normal"
    tie_word_embeddings: bool = True
    upcast_logits_for_loss: bool = False
    # MoE (moe_dolomite/config.py:4-83)
    num_experts: int = 0
    num_experts_per_tok: int = 0
    rope_scaling: dict | None = None  # YaRN: {"factor": s, "original_max_position_embeddings": n}
    # dropout probabilities (config.py:6-111); only applied while a DropoutOracle is installed in DROPOUT (training mode)
    resid_pdrop: float = 0.0
    embd_pdrop: float = 0.0
    attn_pdrop: float = 0.0
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.n_inner is None:
            self.n_inner = 4 * self.n_embd  # config.py:56
        if self.attention_head_type == "mha":
            self.num_key_value_heads = self.n_head if self.num_key_value_heads is None else self.num_key_value_heads
            assert self.num_key_value_heads == self.n_head
        elif self.attention_head_type == "mqa":
            self.num_key_value_heads = 1 if self.num_key_value_heads is None else self.num_key_value_heads
            assert self.num_key_value_heads == 1
        else:
            assert self.num_key_value_heads is not None and self.n_head % self.num_key_value_heads == 0

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head

    @property
    def is_glu(self) -> bool:  # activations/glu.py:49-50
        return self.activation_function.endswith("glu")


def _r(x: torch.Tensor, bf16: bool) -> torch.Tensor:
    """round to bf16 and back when emulating the mixed-precision path"""
    return x.to(torch.bfloat16).to(torch.float32) if bf16 else x


# ------------------------------------------------------------------------------------------------
# dropout masks -- integer arithmetic, bit exact with the kernels (csrc/common.cuh: lowbias32, dropout_hash_flat,
# dropout_hash_qk; kernels.dropout_keys).  The reference draws its masks from torch's Philox stream (nn.Dropout,
# flash-attn's dropout_p), which no independent implementation reproduces; what is restated is the distribution
# (independent Bernoulli(1 - p) keeps, kept values scaled by 1 / (1 - p)) and where the masks enter the arithmetic.
# ------------------------------------------------------------------------------------------------
_U32 = np.uint32


def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(_U32)
    x ^= x >> _U32(16)
    x *= _U32(0x21F0AAAD)
    x ^= x >> _U32(15)
    x *= _U32(0x735A2D97)
    x ^= x >> _U32(15)
    return x


class DropoutOracle:
    """masks of ONE training pass with seed `seed` (the engine's `_dropout_now`); call sites: 0 = embeddings,
    4 i + 1 / + 2 / + 3 = block i attention residual / MLP residual / attention probabilities"""

    def __init__(self, seed: int):
        self.seed = int(seed)

    def keys(self, site: int) -> tuple[int, int]:
        m = (1 << 64) - 1
        z = (self.seed * 0x9E3779B97F4A7C15 + (int(site) + 1) * 0xD1B54A32D192ED03) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        z ^= z >> 31
        return int(z & 0xFFFFFFFF), int(z >> 32)

    @staticmethod
    def threshold(p: float) -> int:
        return int(min(math.floor(float(np.float32(p)) * 4294967296.0 + 0.5), 4294967295))

    def flat_scale(self, site: int, shape, p: float) -> torch.Tensor:
        """1 / (1 - p) where element e (row-major index) is kept, 0 where it is dropped"""
        k0, k1 = self.keys(site)
        with np.errstate(over="ignore"):
            e = np.arange(int(np.prod(shape)), dtype=np.uint64)
            h = _lowbias32(_lowbias32((e & np.uint64(0xFFFFFFFF)).astype(_U32) ^ _U32(k0)) + (e >> np.uint64(32)).astype(_U32) + _U32(k1))
        keep = h >= _U32(self.threshold(p))
        ks = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        return torch.from_numpy(np.where(keep, np.float32(ks), np.float32(0.0)).reshape(shape))

    def attn_scale(self, site: int, head: int, q_tok: np.ndarray, k_tok: np.ndarray, p: float) -> torch.Tensor:
        """[len(q_tok), len(k_tok)] keep scales of one head; q_tok / k_tok are GLOBAL token rows of the packed stream"""
        k0, k1 = self.keys(site)
        with np.errstate(over="ignore"):
            hk = _lowbias32(np.asarray([head], dtype=_U32) * _U32(0xC2B2AE3D) + _U32(k0)) ^ _U32(k1)
            q = np.asarray(q_tok, dtype=_U32)[:, None] * _U32(0x9E3779B1)
            k = np.asarray(k_tok, dtype=_U32)[None, :] * _U32(0x85EBCA77)
            h = _lowbias32(q ^ k ^ hk)
        keep = h >= _U32(self.threshold(p))
        ks = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        return torch.from_numpy(np.where(keep, np.float32(ks), np.float32(0.0)))


# installed by a test for the duration of one training pass (None = evaluation mode: dropout is the identity)
DROPOUT: DropoutOracle | None = None


def _drop(x: torch.Tensor, site: int, p: float, bf16: bool) -> torch.Tensor:
    """nn.Dropout in training mode: bf16(x * mask / (1 - p))"""
    if DROPOUT is None or not p:
        return x
    return _r(x * DROPOUT.flat_scale(site, tuple(x.shape), p), bf16)


# ------------------------------------------------------------------------------------------------
# integer bookkeeping -- bit exact
# ------------------------------------------------------------------------------------------------
def split_tokens(tokens: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """model_wrapper/pretraining.py:191-192: input_ids = tokens[:, :-1]; labels = tokens[:, 1:]"""
    return tokens[:, :-1], tokens[:, 1:]


def prepare_model_inputs(
    input_ids: np.ndarray, eos_token_id: int | None, reset_attention_mask: bool, reset_position_ids: bool
) -> dict:
    """model_wrapper/pretraining.py:129-169 (+ buffers at :201-229) for the padding-free transformer."""
    batch_size, sequence_length = input_ids.shape
    flat = input_ids.reshape(-1)
    if reset_attention_mask:
        ends = flat == eos_token_id
        for i in range(sequence_length - 1, batch_size * sequence_length, sequence_length):
            ends[i] = True
        cu = np.concatenate([[0], np.nonzero(ends)[0] + 1]).astype(np.int32)
        seqlen = cu[1:] - cu[:-1]
        max_seqlen = int(seqlen.max())
        if reset_position_ids:
            pos = np.concatenate([np.arange(0, int(n), 1, dtype=np.int32) for n in seqlen])
        else:
            pos = np.tile(np.arange(sequence_length, dtype=np.int64), batch_size)
    else:
        cu = np.arange(0, batch_size * sequence_length + 1, sequence_length, dtype=np.int32)
        max_seqlen = sequence_length
        pos = np.tile(np.arange(sequence_length, dtype=np.int64), batch_size)
    return {"input_ids": flat, "cu_seqlens": cu, "max_seqlen": max_seqlen, "position_ids": pos}


def convert_padding_free_lists_to_tensors(input_ids: list[list[int]], position_ids=None, labels=None) -> dict:
    """hf_models/utils.py:20-57"""
    seqlens = np.array([0] + [len(x) for x in input_ids], dtype=np.int64)
    cu = np.cumsum(seqlens).astype(np.int32)
    max_seqlen = int(seqlens.max())
    if position_ids is None:
        position_ids = [list(range(len(x))) for x in input_ids]
    flat = lambda ll: np.array([v for row in ll for v in row], dtype=np.int64)  # noqa: E731
    out = {"input_ids": flat(input_ids), "position_ids": flat(position_ids), "cu_seqlens": cu, "max_seqlen": max_seqlen}
    if labels is not None:
        out["labels"] = flat(labels)
    return out


def finetune_shift_labels(labels: np.ndarray, cu_seqlens: np.ndarray) -> np.ndarray:
    """gpt_dolomite/main.py:185-191: shift_labels = labels[1:]; shift_labels[cu_seqlens[1:-1]-1] = -100"""
    shift = labels[1:].copy()
    shift[cu_seqlens[1:-1] - 1] = -100
    return shift


# QKV weight (de)interleave -- attention/utils.py:18-106
def interleave_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    hd, nh, nkv = cfg.head_dim, cfg.n_head, cfg.num_key_value_heads
    if cfg.attention_head_type == "mqa":
        return torch.cat([q, k, v])
    g = nh // nkv
    parts = []
    for i in range(nkv):
        parts.append(q[i * g * hd : (i + 1) * g * hd])
        parts.append(k[i * hd : (i + 1) * hd])
        parts.append(v[i * hd : (i + 1) * hd])
    return torch.cat(parts)


def split_qkv(w: torch.Tensor, cfg: OracleConfig):
    hd, nh, nkv = cfg.head_dim, cfg.n_head, cfg.num_key_value_heads
    if cfg.attention_head_type == "mqa":
        return w.split((nh * hd, hd, hd))
    g = nh // nkv
    x = w.view(nkv, g + 2, hd, *w.shape[1:])
    q = x[:, :g].reshape(-1, *w.shape[1:])
    k = x[:, g].reshape(-1, *w.shape[1:])
    v = x[:, g + 1].reshape(-1, *w.shape[1:])
    return q, k, v


# ------------------------------------------------------------------------------------------------
# float ops
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, bf16: bool = False) -> torch.Tensor:
    """normalization/rmsnorm/base.py:18-25: weight * (x32 * rsqrt(mean(x32^2) + eps)).to(input_dtype)"""
    x32 = x.float()
    var = x32.pow(2).mean(-1, keepdim=True)
    xn = x32 * torch.rsqrt(var + eps)
    return _r(w * _r(xn, bf16), bf16)


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, eps: float, bf16: bool = False) -> torch.Tensor:
    """normalization/layernorm/__init__.py: torch.nn.LayerNorm -- fp32 statistics, a single rounding at the output"""
    return _r(F.layer_norm(x.float(), (x.shape[-1],), w.float(), None if b is None else b.float(), eps), bf16)


def norm(x: torch.Tensor, p: dict, prefix: str, cfg: "OracleConfig", bf16: bool = False) -> torch.Tensor:
    """get_normalization_function (normalization/__init__.py:13-30): rmsnorm | layernorm"""
    if cfg.normalization_function == "rmsnorm":
        return rmsnorm(x, p[prefix + "weight"], cfg.layer_norm_epsilon, bf16)
    if cfg.normalization_function == "layernorm":
        return layernorm(x, p[prefix + "weight"], p.get(prefix + "bias"), cfg.layer_norm_epsilon, bf16)
    raise ValueError(f"oracle: unsupported normalization {cfg.normalization_function}")


def yarn_inv_freq_and_mscale(head_dim: int, base: float, scale: float, original_max_position_embeddings: int,
                             extrapolation_factor: float = 1, attn_factor: float = 1, beta_fast: int = 32, beta_slow: int = 1):
    """YaRNScaledRoPE (position_embedding/rope.py:56-101, :118-145): per-dimension blend of interpolated and extrapolated
    frequencies + the magnitude correction `mscale` that the reference multiplies into the cos / sin tables"""

    def correction_dim(num_rotations):
        return (head_dim * math.log(original_max_position_embeddings / (num_rotations * 2 * math.pi))) / (2 * math.log(base))

    pos_freqs = base ** (torch.arange(0, head_dim, 2).float() / head_dim)
    extrapolation = 1.0 / pos_freqs
    interpolation = 1.0 / (scale * pos_freqs)
    low = max(math.floor(correction_dim(beta_fast)), 0)
    high = min(math.ceil(correction_dim(beta_slow)), head_dim - 1)
    hi = high + 0.001 if low == high else high
    ramp = torch.clamp((torch.arange(head_dim // 2, dtype=torch.float32) - low) / (hi - low), 0, 1)
    mask = (1 - ramp) * extrapolation_factor
    inv_freq = interpolation * (1 - mask) + extrapolation * mask
    mscale = (1.0 if scale <= 1 else 0.1 * math.log(scale) + 1.0) * attn_factor
    return inv_freq, mscale


def rope_tables(head_dim: int, n_positions: int, base: float, bf16: bool = False, rope_scaling: dict | None = None):
    """position_embedding/rope.py:25-55; with `rope_scaling` (YaRN, gpt_dolomite/base.py:534-547) the frequencies and the
    table magnitude change, nothing else"""
    mscale = 1.0
    if rope_scaling is None:
        inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    else:
        inv_freq, mscale = yarn_inv_freq_and_mscale(head_dim, base, rope_scaling["factor"],
                                                    rope_scaling["original_max_position_embeddings"])
    t = torch.arange(n_positions, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return _r(emb.cos() * mscale, bf16), _r(emb.sin() * mscale, bf16)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, bf16: bool = False) -> torch.Tensor:
    """position_embedding/rope.py:104-114: (x*cos) + (rotate_half(x)*sin), a rounding after each op in bf16"""
    x1, x2 = torch.chunk(x, 2, dim=-1)
    rot = torch.cat((-x2, x1), dim=-1)
    return _r(_r(x * cos, bf16) + _r(rot * sin, bf16), bf16)


def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, bf16: bool = False) -> torch.Tensor:
    """modeling_utils/linear.py:5-25 (nn.Linear): fp32 accumulate, one rounding of the biased result"""
    return _r(F.linear(x, w, b), bf16)


def softmax_scale(cfg: OracleConfig) -> float:
    """attention/base.py:277-286"""
    if not cfg.scale_attn_weights:
        return 1.0
    return cfg.attention_multiplier if cfg.attention_multiplier is not None else 1.0 / math.sqrt(cfg.head_dim)


def split_qkv_activations(qkv: torch.Tensor, cfg: OracleConfig):
    """attention/padding_free.py:79-116 -- q [T,nh,hd], k/v [T,nkv,hd] from the packed c_attn output"""
    T = qkv.shape[0]
    hd, nh, nkv = cfg.head_dim, cfg.n_head, cfg.num_key_value_heads
    if cfg.attention_head_type == "mqa":
        q, k, v = qkv.split((nh * hd, hd, hd), dim=-1)
        return q.reshape(T, nh, hd), k.unsqueeze(1), v.unsqueeze(1)
    g = nh // nkv
    x = qkv.view(T, nkv, (g + 2) * hd)
    q, k, v = x.split((g * hd, hd, hd), dim=-1)
    return q.reshape(T, nh, hd), k, v


def packed_causal_attention(q, k, v, cu_seqlens: np.ndarray, scale: float, bf16: bool = False, dropout_site: int | None = None,
                            dropout_p: float = 0.0) -> torch.Tensor:
    """Block-diagonal (per document) causal softmax attention = flash_attn_varlen_func(causal=True) at
    attention/padding_free.py:51-62; arithmetic follows the eager Attention (attention/base.py:171-275) with fp32
    softmax (attention_softmax_in_fp32).  q [T,nh,hd]; k,v [T,nkv,hd].  Returns [T, nh*hd]."""
    T, nh, hd = q.shape
    nkv = k.shape[1]
    g = nh // nkv
    k = k.repeat_interleave(g, dim=1) if g > 1 else k
    v = v.repeat_interleave(g, dim=1) if g > 1 else v
    out = torch.zeros(T, nh, hd, dtype=torch.float32)
    for d in range(len(cu_seqlens) - 1):
        s, e = int(cu_seqlens[d]), int(cu_seqlens[d + 1])
        if e == s:
            continue
        qd, kd, vd = q[s:e].transpose(0, 1), k[s:e].transpose(0, 1), v[s:e].transpose(0, 1)
        sc = torch.matmul(qd, kd.transpose(1, 2)) * scale
        mask = torch.ones(e - s, e - s, dtype=torch.bool).tril()
        sc = sc.masked_fill(~mask, float("-inf"))
        p = torch.softmax(sc.float(), dim=-1)
        if DROPOUT is not None and dropout_p and dropout_site is not None:
            # attention-probability dropout (attention/base.py:252 attn_dropout; flash_attn_varlen_func(dropout_p) at
            # padding_free.py:49-59): the softmax normaliser is that of the UNdropped row
            tok = np.arange(s, e)
            p = p * torch.stack([DROPOUT.attn_scale(dropout_site, h, tok, tok, dropout_p) for h in range(nh)])
        out[s:e] = torch.matmul(_r(p, bf16), vd).transpose(0, 1)
    return _r(out.reshape(T, nh * hd), bf16)


def activation(x: torch.Tensor, name: str, bf16: bool = False) -> torch.Tensor:
    """activations/glu.py:26-28 (x.chunk(2)[0] * act(x.chunk(2)[1])) and activations/base.py"""
    if name == "swiglu":
        u, g = x.chunk(2, dim=-1)
        return _r(u * _r(F.silu(g), bf16), bf16)
    if name == "gelu_pytorch_tanh":
        return _r(F.gelu(x, approximate="tanh"), bf16)
    raise ValueError(f"oracle: unsupported activation {name}")


def mlp(x, p: dict, prefix: str, cfg: OracleConfig, bf16: bool = False) -> torch.Tensor:
    """gpt_dolomite/mlp.py:45-50"""
    h = linear(x, p[prefix + "c_fc.weight"], p.get(prefix + "c_fc.bias"), bf16)
    h = activation(h, cfg.activation_function, bf16)
    return linear(h, p[prefix + "c_proj.weight"], p.get(prefix + "c_proj.bias"), bf16)


# ---- MoE (moe_dolomite/moe/base.py:108-181) ----
def moe_route(x, gate_w, top_k: int, bf16: bool = False):
    """_compute_routing_weights / _get_topk: topk on raw logits, fp32 softmax over the selected k"""
    logits = linear(x, gate_w, None, bf16)
    if top_k == 1:
        w, idx = logits.max(dim=-1, keepdim=True)
    else:
        w, idx = logits.topk(top_k, dim=-1)
    w = _r(torch.softmax(w.float(), dim=-1), bf16)
    return w, idx, logits


def moe_expert_counts(idx: torch.Tensor, num_experts: int) -> np.ndarray:
    """moe/base.py:158-164: bincount(selected_experts.flatten(), minlength=E)  (bit exact)"""
    return np.bincount(idx.reshape(-1).numpy(), minlength=num_experts)


# Tests may pin the expert choice per layer prefix ({"transformer.h.0.mlp.": LongTensor[T,k]}) so that a comparison
# against a lower-precision implementation is not dominated by top-k flips of near-tied router logits; the routing
# WEIGHTS are still recomputed here from the oracle's own logits.
FORCED_ROUTING: dict = {}


def sparse_moe(x, p: dict, prefix: str, cfg: OracleConfig, bf16: bool = False):
    """SparseMoE.forward (moe/base.py:108-173): per expert linear -> act -> linear, gate-weighted index_add"""
    T, H = x.shape
    w, idx, logits = moe_route(x, p[prefix + "gate.weight"], cfg.num_experts_per_tok, bf16)
    if prefix in FORCED_ROUTING:
        idx = FORCED_ROUTING[prefix]
        w = _r(torch.softmax(logits.gather(1, idx).float(), dim=-1), bf16)
    out = torch.zeros(T, H, dtype=torch.float32)
    Wfc, Wproj = p[prefix + "c_fc.weight"], p[prefix + "c_proj.weight"]  # [E, out, in]
    bfc, bproj = p.get(prefix + "c_fc.bias"), p.get(prefix + "c_proj.bias")
    for e in range(cfg.num_experts):
        tok, slot = torch.nonzero(idx == e, as_tuple=True)
        if tok.numel() == 0:
            continue
        h = linear(x[tok], Wfc[e], None if bfc is None else bfc[e], bf16)
        h = activation(h, cfg.activation_function, bf16)
        h = linear(h, Wproj[e], None if bproj is None else bproj[e], bf16)
        out.index_add_(0, tok, _r(h * w[tok, slot].unsqueeze(-1), bf16))
    return _r(out, bf16), logits


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------
def block(h, p: dict, i: int, cfg: OracleConfig, cos, sin, cu_seqlens, bf16: bool = False) -> torch.Tensor:
    """GPTDolomiteBlock.forward (gpt_dolomite/layer.py:49-87) with PaddingFreeAttention
    (attention/padding_free.py:15-77); MoE block: moe_dolomite/layer.py:51-95."""
    pre = f"transformer.h.{i}."
    res = h
    x = norm(h, p, pre + "ln_1.", cfg, bf16)
    qkv = linear(x, p[pre + "attn.c_attn.weight"], p.get(pre + "attn.c_attn.bias"), bf16)
    q, k, v = split_qkv_activations(qkv, cfg)
    if cfg.position_embedding_type == "rope":
        q = apply_rope(q, cos, sin, bf16)
        k = apply_rope(k, cos, sin, bf16)
    a = packed_causal_attention(q, k, v, cu_seqlens, softmax_scale(cfg), bf16, dropout_site=4 * i + 3, dropout_p=cfg.attn_pdrop)
    a = linear(a, p[pre + "attn.c_proj.weight"], p.get(pre + "attn.c_proj.bias"), bf16)
    a = _drop(a, 4 * i + 1, cfg.resid_pdrop, bf16)  # resid_dropout (attention/padding_free.py:75)
    if cfg.m_residual is not None:
        a = _r(a * cfg.m_residual, bf16)
    h = _r(a + res, bf16)
    res = h
    x = norm(h, p, pre + "ln_2.", cfg, bf16)
    if cfg.num_experts > 0:
        m, _ = sparse_moe(x, p, pre + "mlp.", cfg, bf16)
    else:
        m = mlp(x, p, pre + "mlp.", cfg, bf16)
    m = _drop(m, 4 * i + 2, cfg.resid_pdrop, bf16)  # gpt_dolomite/mlp.py:49, moe/base.py:120
    if cfg.m_residual is not None:
        m = _r(m * cfg.m_residual, bf16)
    return _r(res + m, bf16)


def forward_logits(p: dict, cfg: OracleConfig, input_ids, position_ids, cu_seqlens, bf16: bool = False, return_hidden=False):
    """GPTDolomiteModel.forward + get_lm_logits (gpt_dolomite/base.py:170-244, :351-372, :289-296;
    gpt_dolomite/main.py:143-177).  `p`: reference state-dict names -> fp32 tensors (rounded to bf16 first when
    emulating FSDP param_dtype=bf16)."""
    if bf16:
        p = {k: _r(v, True) for k, v in p.items()}
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    pos = torch.as_tensor(np.asarray(position_ids), dtype=torch.long)
    h = p["transformer.wte.weight"][ids]
    if cfg.position_embedding_type == "learned_absolute":  # gpt_dolomite/base.py:351-372: wte(ids) + wpe(position_ids)
        h = _r(h + p["transformer.wpe.weight"][pos], bf16)
    h = _drop(h, 0, cfg.embd_pdrop, bf16)  # gpt_dolomite/base.py:368 `self.drop`
    if cfg.m_emb is not None:
        h = _r(h * cfg.m_emb, bf16)
    cos = sin = None
    if cfg.position_embedding_type == "rope":
        ct, st = rope_tables(cfg.head_dim, cfg.n_positions, cfg.rope_theta, bf16, cfg.rope_scaling)
        cos, sin = ct[pos].unsqueeze(1), st[pos].unsqueeze(1)
    hidden = [h]
    for i in range(cfg.n_layer):
        h = block(h, p, i, cfg, cos, sin, cu_seqlens, bf16)
        hidden.append(h)
    h = norm(h, p, "transformer.ln_f.", cfg, bf16)
    head = p["transformer.wte.weight"] if cfg.tie_word_embeddings else p["lm_head.weight"]
    logits = linear(h, head, None, bf16)
    if cfg.m_width is not None:
        logits = _r(logits / cfg.m_width, bf16)
    return (logits, hidden) if return_hidden else logits


def pretraining_loss(p: dict, cfg: OracleConfig, tokens: np.ndarray, eos_token_id=None, reset_attention_mask=False,
                     reset_position_ids=False, bf16: bool = False):
    """ModelWrapperForPretraining.forward (model_wrapper/pretraining.py:89-127): mean CE over all T positions of
    labels = tokens[:, 1:].  Returns (loss, logits)."""
    input_ids, labels = split_tokens(np.asarray(tokens))
    b = prepare_model_inputs(input_ids, eos_token_id, reset_attention_mask, reset_position_ids)
    logits = forward_logits(p, cfg, b["input_ids"], b["position_ids"], b["cu_seqlens"], bf16)
    lab = torch.as_tensor(np.ascontiguousarray(labels).reshape(-1), dtype=torch.long)
    loss = F.cross_entropy(logits.float(), lab)
    return loss, logits


def finetuning_loss(p: dict, cfg: OracleConfig, input_ids: list[list[int]], labels: list[list[int]], bf16=False):
    """padding-free finetune path: hf_models/utils.py:20-57 + gpt_dolomite/main.py:179-202"""
    b = convert_padding_free_lists_to_tensors(input_ids, labels=labels)
    logits = forward_logits(p, cfg, b["input_ids"], b["position_ids"], b["cu_seqlens"], bf16)
    shift = finetune_shift_labels(b["labels"], b["cu_seqlens"])
    loss = F.cross_entropy(logits[:-1].float(), torch.as_tensor(shift, dtype=torch.long), ignore_index=-100)
    return loss, logits


# ------------------------------------------------------------------------------------------------
# parameter init (linear.py:18-25, embedding.py:36-44, attention/base.py:73-86, mlp.py:26-41, moe/base.py:75-104)
# ------------------------------------------------------------------------------------------------
def init_params(cfg: OracleConfig, seed: int = 42) -> dict:
    g = torch.Generator().manual_seed(seed)
    H, F_, V, L = cfg.n_embd, cfg.n_inner, cfg.vocab_size, cfg.n_layer
    std = cfg.initializer_range
    if cfg.init_method == "mup":
        std /= math.sqrt(cfg.m_width)
    std_proj = cfg.initializer_range / math.sqrt(2 * L)
    if cfg.init_method == "mup":
        std_proj /= math.sqrt(cfg.m_width)
    n = lambda *s, sd: torch.randn(*s, generator=g) * sd  # noqa: E731
    p = {"transformer.wte.weight": n(V, H, sd=cfg.initializer_range)}
    qkv_out = H + 2 * cfg.num_key_value_heads * cfg.head_dim
    fc_out = 2 * F_ if cfg.is_glu else F_
    for i in range(L):
        pre = f"transformer.h.{i}."
        p[pre + "ln_1.weight"] = torch.ones(H)
        p[pre + "ln_2.weight"] = torch.ones(H)
        p[pre + "attn.c_attn.weight"] = n(qkv_out, H, sd=std)
        p[pre + "attn.c_proj.weight"] = n(H, H, sd=std_proj)
        if cfg.num_experts > 0:
            E = cfg.num_experts
            p[pre + "mlp.gate.weight"] = n(E, H, sd=std)
            p[pre + "mlp.c_fc.weight"] = n(E, fc_out, H, sd=std)
            p[pre + "mlp.c_proj.weight"] = n(E, H, F_, sd=std_proj)
        else:
            p[pre + "mlp.c_fc.weight"] = n(fc_out, H, sd=std)
            p[pre + "mlp.c_proj.weight"] = n(H, F_, sd=std_proj)
        if cfg.add_bias:
            p[pre + "attn.c_attn.bias"] = torch.zeros(qkv_out)
            p[pre + "attn.c_proj.bias"] = torch.zeros(H)
            if cfg.num_experts > 0:
                p[pre + "mlp.c_fc.bias"] = torch.zeros(cfg.num_experts, fc_out)
                p[pre + "mlp.c_proj.bias"] = torch.zeros(cfg.num_experts, H)
            else:
                p[pre + "mlp.c_fc.bias"] = torch.zeros(fc_out)
                p[pre + "mlp.c_proj.bias"] = torch.zeros(H)
    p["transformer.ln_f.weight"] = torch.ones(H)
    if not cfg.tie_word_embeddings:
        p["lm_head.weight"] = n(V, H, sd=cfg.initializer_range)  # main.py:19-21: no muP width scaling on the head
    # drawn last so that the random stream of every earlier configuration is unchanged
    if cfg.position_embedding_type == "learned_absolute":
        p["transformer.wpe.weight"] = n(cfg.n_positions, H, sd=cfg.initializer_range)
    if cfg.normalization_function == "layernorm":  # nn.LayerNorm always carries a bias (zeros at init)
        for i in range(L):
            p[f"transformer.h.{i}.ln_1.bias"] = torch.zeros(H)
            p[f"transformer.h.{i}.ln_2.bias"] = torch.zeros(H)
        p["transformer.ln_f.bias"] = torch.zeros(H)
    return p


def model_flops_per_token(cfg: OracleConfig, seq_len: int) -> float:
    """train_utils.py:197-236 (per token, no checkpointing): 3*L*[4h(h(1+k/n)+s) + (6|4)hf] + 6hv"""
    h, f, n, k, L, v = cfg.n_embd, cfg.n_inner, cfg.n_head, cfg.num_key_value_heads, cfg.n_layer, cfg.vocab_size
    mlp_f = (6 if cfg.is_glu else 4) * h * f
    if cfg.num_experts > 0:
        mlp_f = mlp_f * cfg.num_experts_per_tok + 2 * h * cfg.num_experts
    attn = 4 * h * (h * (1 + k / n) + seq_len)
    return 3 * L * (attn + mlp_f) + 6 * h * v
